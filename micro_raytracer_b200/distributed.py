"""Multi-GPU rendering: sample split + one reduce of the accumulation buffers.

The path shards by samples: rank r of G renders the global sample indices
r, r+G, r+2G, ... of every supersampled pixel, then the per-rank accumulators are summed onto
`dst` — the only exchange step of the path (it replaces the `Mutex<HashMap>` merge of
src/sampler.rs:60-70).  One process per GPU; `torch.distributed` is the plumbing (NCCL over
NVLink for device accumulators; any backend for host accumulators, which is what the
world-size-2 gloo test on CPU drives).
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def passes_of_rank(spp: int, rank: int, world: int) -> int:
    """How many of the global samples 0..spp-1 rank `rank` renders (indices rank, rank+world, ...)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad partition rank {rank} of {world}")
    return len(range(rank, spp, world))


def reduce_accum(sampler, total_passes: int, group=None, dst: int = 0, device_tensor=None) -> Optional[np.ndarray]:
    """Sum the samplers' accumulators onto rank `dst` and record there that the buffer now
    holds `total_passes` passes.

    device_tensor: a torch CUDA tensor aliasing the sampler's device accumulator
    (torch.as_tensor(sampler.accum_device()[0], device=...)); the reduce then runs in place on the device.
    Ordering: when the sampler was bound to torch's current stream (sampler.set_stream(stream.cuda_stream),
    what bench.py does) the path kernels, the reduce and the later film kernels are ordered by that one
    stream and nothing waits on the host.  Otherwise the sampler's kernels run on the context's private
    stream, which torch knows nothing about: the sampler is synchronised before the reduce and torch's
    stream after it, so the reduce never reads a film that is still being written and img() never
    tone-maps one the reduce has not finished.
    Without a device tensor the host accumulator (sampler.accum()) is reduced and the sum is returned on
    `dst` (None elsewhere)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if device_tensor is not None:
        cur = torch.cuda.current_stream(device_tensor.device)
        shared = sampler.stream is not None and sampler.stream == cur.cuda_stream
        sampler.accum_device()          # launches whatever passes are still queued in the library
        if not shared:
            sampler.sync()
        if world > 1:
            dist.reduce(device_tensor, dst=dst, op=dist.ReduceOp.SUM, group=group)
        if not shared:
            cur.synchronize()
        sampler.set_passes(total_passes)
        return None
    acc, _ = sampler.accum()
    t = torch.from_numpy(np.ascontiguousarray(acc))
    if world > 1:
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return t.numpy() if rank == dst else None


def render_distributed(sampler, scene, frame, rt, spp: int, group=None, dst: int = 0, device_tensor=None):
    """Render `spp` passes split over the ranks of `group` and reduce onto `dst`.
    Returns the summed host accumulator on `dst` when no device tensor is given."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    sampler._bind(scene, frame, rt)
    sampler.reset()
    sampler.set_partition(rank, world)
    n = passes_of_rank(spp, rank, world)
    if device_tensor is not None:
        sampler.execute_async(n)
    elif n:
        sampler.execute(scene, frame, rt, n)
    return reduce_accum(sampler, spp, group, dst, device_tensor)


class FilmGather:
    """Film read-out of a multi-process render WITHOUT reducing the accumulators: every rank tone-maps one band of
    pixels, reading that band of every rank's accumulator over CUDA IPC peer mappings (NVLink), and writes the u8 pixels
    into rank 0's supersampled image; rank 0 then resizes and copies out (include/mrt.h: mrt_ipc_*).  Against
    reduce_accum + img() this moves 1/world of the 16-byte-per-pixel film per rank instead of all of it onto one rank,
    and fuses the exchange into the tone-map kernel.

    The sampler must share torch's current stream (sampler.set_stream(stream.cuda_stream)): the two device-side
    barriers are one-element NCCL all-reduces queued on that stream, so nothing waits on the host."""

    def __init__(self, sampler, group=None):
        import torch
        import torch.distributed as dist
        self.s, self.group = sampler, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        self._token = torch.zeros(1, device=dev)
        handles = [None] * self.world
        dist.all_gather_object(handles, sampler.ipc_export(), group=group)
        sampler.ipc_attach(self.rank, self.world, [h[0] for h in handles], handles[0][1])
        dist.barrier(group=group)  # nobody renders into a buffer a peer has not mapped yet

    def _device_barrier(self):
        import torch.distributed as dist
        dist.all_reduce(self._token, group=self.group)

    def bands(self, total_passes: int):
        """Queue the gather: barrier, this rank's band, barrier.  Returns at once (everything is stream-ordered)."""
        self.s.accum_device()          # launches whatever passes are still queued in the library
        self._device_barrier()         # every rank's passes are rendered ...
        self.s.ipc_tonemap_band(total_passes)
        self._device_barrier()         # ... and every band is written

    def img(self, total_passes: int) -> Optional[np.ndarray]:
        """≙ Sampler::img of the whole job: the (res_h, res_w, 3) uint8 image on rank 0, None elsewhere."""
        self.bands(total_passes)
        return self.s.img_gathered() if self.rank == 0 else None
