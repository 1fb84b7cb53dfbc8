"""Multi-GPU plumbing of the hot path: one process per GPU, the scene replicated on
every rank, the points / pixels of ONE job split over the ranks.

  * IrT generation (forward only, models/tracer_o3d_irt.py:156-173): the texels are taken along a Morton curve and
    dealt out to the ranks in blocks of 512 (block-cyclic, so every rank gets the same mix of cheap and expensive
    regions); each rank's kernel stores its results straight into EVERY rank's result buffer over NVLink peer
    memory (PeerGroup, texir_irradiance_fwd_peers) -- the all-gather is the compute kernel's epilogue -- followed
    by one flag barrier.  `collective="nccl"` is the library formulation of the same step (zeroed buffer + one
    all-reduce of [P,3]) for comparison and for backends without peer mapping (gloo in the CPU tests).
  * Material stage (forward + backward, trainer/train_material.py:448-450): the pixels of the cube map are dealt out
    in blocks of 128 (the G-buffer kernel's CTA), every rank scatters its pixels' gradients into its own copy of the
    texture-gradient buffers, and ONE all-reduce(sum) runs over the flat buffer both gradients live in
    (ops._MatStage allocates them as one tensor, so there is no concatenation and no copy back).

Works with any torch.distributed backend for the NCCL-style paths; the peer path needs CUDA IPC between the ranks.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


# ---------------------------------------------------------------------------------------------------
# who computes what
# ---------------------------------------------------------------------------------------------------
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


_cyclic_cache = {}


def block_cyclic_index(num: int, rank: int, world: int, block: int, device=None) -> torch.Tensor:
    """int64 positions j in [0, num) with (j // block) % world == rank, ascending: the share of rank `rank` when
    `num` items are dealt out block by block"""
    if world <= 0 or not (0 <= rank < world) or block <= 0:
        raise ValueError("bad rank / world / block")
    key = (num, rank, world, block, str(device))
    hit = _cyclic_cache.get(key)
    if hit is None:
        j = torch.arange(num, dtype=torch.int64, device=device)
        hit = j[(j // block) % world == rank]
        if len(_cyclic_cache) > 64:
            _cyclic_cache.clear()
        _cyclic_cache[key] = hit
    return hit


def local_pixels(full: torch.Tensor, rank: int, world: int, views: int, block: int = 128) -> torch.Tensor:
    """the pixels of a [views, h, w, C] image (e.g. the ground truth of a cube map) that rank `rank` renders in a
    sharded Mat-stage step, flat [views*h*w/world, C] in the order ops.gbuffer_mvp(shard=(rank, world)) produces:
    128-pixel blocks rank, rank + world, ... of every view"""
    C = full.shape[-1]
    per_view = full.reshape(views, -1, C)
    n = per_view.shape[1]
    if world == 1:
        return per_view.reshape(-1, C)
    if n % (block * world) != 0:
        raise ValueError("pixels per view must be a multiple of block * world")
    sel = block_cyclic_index(n, rank, world, block, full.device)
    return per_view[:, sel, :].reshape(-1, C)


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


# ---------------------------------------------------------------------------------------------------
# library collectives
# ---------------------------------------------------------------------------------------------------
def gather_points(local: torch.Tensor, num_points: int, align: int = 1, group=None) -> torch.Tensor:
    """all_gather of per-rank CONTIGUOUS slices [n_r, C] (ragged allowed) into the full [num_points, C] tensor"""
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


def all_gather_rows(rows: torch.Tensor, group=None) -> torch.Tensor:
    """[n, C] on every rank (same n) -> [world * n, C], rank-major; one collective"""
    world = dist.get_world_size(group)
    out = torch.empty((world * rows.shape[0],) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, rows.contiguous(), group=group)
    else:  # gloo (the CPU-side tests): list form
        dist.all_gather(list(out.chunk(world, 0)), rows.contiguous(), group=group)
    return out


def _flat_view_of(grads: List[torch.Tensor]) -> Optional[torch.Tensor]:
    """the ONE contiguous buffer the gradients tile, if they do (ops._MatStage returns both texture gradients as
    views of one allocation and autograd hands those views out as .grad unchanged); None otherwise"""
    if not grads or any(not g.is_contiguous() or g.dtype != grads[0].dtype or g.device != grads[0].device for g in grads):
        return None
    st = grads[0].untyped_storage()
    if any(g.untyped_storage().data_ptr() != st.data_ptr() for g in grads):
        return None
    spans = sorted((g.storage_offset(), g.numel()) for g in grads)
    off = spans[0][0]
    for o, n in spans:
        if o != off:
            return None
        off += n
    return torch.empty(0, dtype=grads[0].dtype, device=grads[0].device).set_(st, spans[0][0], (off - spans[0][0],))


def allreduce_gradients(params: Iterable[torch.Tensor], group=None, average: bool = False) -> int:
    """ONE all-reduce over the gradients of `params` (materials_a 2048^2x3 + materials_r 4096^2x1 = 112 MiB fp32 at
    the reference's sizes, mat_nvdiffrast.py:68-69).  When the gradients already share one buffer (the Mat-stage
    backward of this package) it is reduced in place; otherwise they are packed, reduced and copied back.
    Returns the number of bytes reduced."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    flat = _flat_view_of(grads)
    packed = flat is None
    if packed:
        flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    if packed:
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
    return flat.numel() * flat.element_size()


# ---------------------------------------------------------------------------------------------------
# peer memory (CUDA IPC over NVLink / NVSwitch): buffers every rank of the node can store into
# ---------------------------------------------------------------------------------------------------
class _DevArray:
    """raw device memory as a __cuda_array_interface__ object, so that torch can view it without owning it"""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerBuffer:
    """`local`: this rank's buffer as a tensor; `ptrs`: the same buffer on every rank (own pointer at index rank)"""

    def __init__(self, group: "PeerGroup", shape, dtype=torch.float32):
        self.group = group
        n = 1
        for s in shape:
            n *= int(s)
        self.nbytes = max(n * torch.empty(0, dtype=dtype).element_size(), 16)
        lib = _lib.load()
        dev = group.device.index
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(lib.texir_peer_alloc(dev, self.nbytes, ctypes.byref(ptr), handle), "texir_peer_alloc")
        self._own = ptr.value
        handles = [None] * group.world
        dist.all_gather_object(handles, bytes(handle), group=group.group)
        self.ptrs: List[int] = []
        self._opened: List[int] = []
        for r, h in enumerate(handles):
            if r == group.rank:
                self.ptrs.append(self._own)
                continue
            p = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            _lib.check(lib.texir_peer_open(dev, buf, ctypes.byref(p)), "texir_peer_open")
            self.ptrs.append(p.value)
            self._opened.append(p.value)
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
        with torch.cuda.device(group.device):
            self.local = torch.as_tensor(_DevArray(self._own, shape, typestr), device=group.device)
        self.ptr_array = (ctypes.c_void_p * group.world)(*self.ptrs)

    def close(self):
        lib = _lib.load()
        dev = self.group.device.index
        for p in self._opened:
            lib.texir_peer_close(dev, ctypes.c_void_p(p))
        self._opened = []
        if self._own:
            lib.texir_peer_free(dev, ctypes.c_void_p(self._own))
            self._own = 0


class PeerGroup:
    """the ranks of one node with their flag arrays mapped into each other (one process per GPU)"""

    def __init__(self, device, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise _lib.TexirError("PeerGroup needs an initialised torch.distributed process group")
        self.device = torch.device(device)
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise _lib.TexirError("peer memory spans the GPUs of one node (at most 8 ranks)")
        self.epoch = 0
        self._buffers: List[PeerBuffer] = []
        self.flags = self.buffer((self.world + 1,), torch.int32)
        dist.barrier(group=group)  # every rank has mapped every flag array before the first device barrier

    def buffer(self, shape, dtype=torch.float32) -> PeerBuffer:
        b = PeerBuffer(self, shape, dtype)
        self._buffers.append(b)
        return b

    def barrier(self) -> None:
        """device-side flag barrier on the current stream: stores issued by earlier kernels of every rank's stream are
        visible to all ranks once it has passed (texir_peer_barrier)"""
        self.epoch += 1
        with torch.cuda.device(self.device):
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(_lib.load().texir_peer_barrier(self.flags.ptr_array, self.rank, self.world, self.epoch,
                                                      self.device.index, st), "texir_peer_barrier")

    def timed_out(self) -> bool:
        """True if a barrier gave up waiting for a rank (synchronises the stream)"""
        return int(self.flags.local[self.world].item()) != 0

    def close(self):
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # nobody unmaps while a peer may still store
        for b in self._buffers:
            b.close()
        self._buffers = []


# ---------------------------------------------------------------------------------------------------
# sharded IrT
# ---------------------------------------------------------------------------------------------------
class ShardedIrradiance:
    """ONE IrT job over all ranks.  `position`, `normal`, `shift` are the FULL [P,.] tensors (identical on every
    rank, so that the shift stream is the single-GPU one); every rank ends up with the full [P,3] result.

        collective="peer"  results are stored into every rank's buffer by the kernel itself (NVLink peer stores),
                           then one flag barrier                                     -- no library collective
        collective="nccl"  results go into a zeroed local buffer, then one all-reduce(sum) of [P,3]
    """

    def __init__(self, tracer, num_points: int, collective: str = "peer", block: int = 512, group=None):
        self.tracer, self.P, self.block, self.group = tracer, int(num_points), int(block), group
        self.rank, self.world = _world(group)
        self.device = tracer.scene.device
        if collective not in ("peer", "nccl"):
            raise ValueError("collective must be 'peer' or 'nccl'")
        self.collective = collective if self.world > 1 else "none"
        self.kernel_events = None
        self.peers = None
        if self.collective == "peer":
            self.peers = PeerGroup(self.device, group)
            self.out = self.peers.buffer((self.P, 3), torch.float32)
            self.result = self.out.local
        else:
            self.result = torch.zeros(self.P, 3, dtype=torch.float32, device=self.device)

    def __call__(self, position: torch.Tensor, normal: torch.Tensor, shift: torch.Tensor) -> torch.Tensor:
        from . import ops
        from .models import _upload_texture
        scene = self.tracer.scene
        _upload_texture(self.tracer, scene)
        N, mode = int(self.tracer.sample_l[0]), self.tracer.sample_type[0]
        # Morton order of ALL texels (the same on every rank); a rank takes every world-th block of it
        order = ops.point_order(scene, position) if (self.world > 1 or self.P > 4096) else None
        mine = order
        if self.world > 1:
            mine = order[block_cyclic_index(self.P, self.rank, self.world, self.block, order.device)]
        ev = None
        if self.kernel_events is not None:   # bench.py: CUDA events around the dominant kernel's launch alone
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            if self.collective == "nccl":
                self.result.zero_()
            ev[0].record(torch.cuda.current_stream(self.device))
        elif self.collective == "nccl":
            self.result.zero_()
        if self.collective == "peer":
            ops.irradiance_peers(scene, position, normal, shift, mine, N, mode, self.out.ptr_array, self.world)
        else:
            ops.irradiance(scene, position, normal, shift, N, mode, out=self.result, order=mine, coherent=False)
        if ev is not None:
            ev[1].record(torch.cuda.current_stream(self.device))
            self.kernel_events.append(ev)
        if self.collective == "peer":
            self.peers.barrier()
        elif self.collective == "nccl":
            dist.all_reduce(self.result, op=dist.ReduceOp.SUM, group=self.group)
        return self.result

    def close(self):
        if self.peers is not None:
            self.peers.close()
            self.peers = None


def sharded_irradiance(tracer, position: torch.Tensor, normal: torch.Tensor, shift: Optional[torch.Tensor] = None,
                       group=None, align: int = 512) -> torch.Tensor:
    """one-shot form (library collective): contiguous 512-aligned slices, one all_gather; returns the full [P,3]"""
    P = position.shape[0]
    rank, world = _world(group)
    b, e = shard_bounds(P, rank, world, align)
    if shift is None:
        from .sample_util import draw_shift
        shift = draw_shift(P, position.device)
    local = tracer.irradiance(position[b:e], normal[b:e], shift[b:e])
    return gather_points(local, P, align, group)
