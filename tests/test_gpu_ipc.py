"""Film gather across PROCESSES (one rank per GPU, the torchrun shape): every rank tone-maps one band of pixels reading
every rank's accumulator over CUDA IPC peer mappings and writes into rank 0's image (include/mrt.h: mrt_ipc_*,
distributed.FilmGather).  Must give the one-GPU image up to the f32 summation order, and the NCCL-reduce path's."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_devices():
    import torch
    return torch.cuda.device_count()


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist

    import micro_raytracer_b200 as mrt
    from micro_raytracer_b200.distributed import FilmGather, passes_of_rank, reduce_accum
    from util import load
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        r = load("CornellBox2", (101, 67), 2.0)
        spp = 2 * world + 1
        s = mrt.Sampler(device=rank)
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        s.set_stream(stream.cuda_stream)
        s._bind(r.scene, r.frame, r.rt)
        s.set_partition(rank, world)
        gather = FilmGather(s)
        imgs = []
        for rep in range(2):                      # twice: the mappings are reused, the films restart
            s.reset()
            for _ in range(passes_of_rank(spp, rank, world)):
                s.execute(r.scene, r.frame, r.rt)  # the reference's loop on every rank
            imgs.append(gather.img(spp))
        # the NCCL-reduce path on the same films, for comparison
        acc = torch.as_tensor(s.accum_device()[0], device=f"cuda:{rank}")
        reduce_accum(s, spp, device_tensor=acc)
        if rank == 0:
            assert np.array_equal(imgs[0], imgs[1])
            np.savez(out_path, gathered=imgs[0], reduced=s.img(r.frame))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif("_n_devices() < 2", reason="needs two GPUs")
def test_ipc_film_gather_equals_the_single_gpu_image(tmp_path):
    import torch.multiprocessing as mp

    import micro_raytracer_b200 as mrt
    from util import load
    world = min(_n_devices(), 4)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "imgs.npz")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = np.load(out)
    r = load("CornellBox2", (101, 67), 2.0)
    one = mrt.Sampler(device=0)
    one.execute(r.scene, r.frame, r.rt, 2 * world + 1)
    want = one.img(r.frame).astype(int)
    for k in ("gathered", "reduced"):
        d = np.abs(got[k].astype(int) - want)
        assert d.max() <= 1 and (d == 0).mean() > 0.999, k
    d = np.abs(got["gathered"].astype(int) - got["reduced"].astype(int))
    assert d.max() <= 1 and (d == 0).mean() > 0.999


def test_ipc_entry_points_fail_cleanly_without_a_gather():
    import micro_raytracer_b200 as mrt
    from micro_raytracer_b200.sampler import MrtError
    from util import load
    r = load("Default", (32, 18), 1.0)
    s = mrt.Sampler(device=0)
    with pytest.raises(MrtError):
        s.ipc_export()                      # no frame yet
    s.execute(r.scene, r.frame, r.rt, 1)
    with pytest.raises(MrtError):
        s.ipc_tonemap_band(1)               # not attached
    with pytest.raises(MrtError):
        s.img_gathered()
    a, i = s.ipc_export()
    assert len(a) == len(i) == 64 and a != i
    with pytest.raises(MrtError):
        s.ipc_attach(2, 2, [a, a], i)       # rank out of range
