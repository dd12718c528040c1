"""World-size-2 gloo test of the multi-GPU host logic (sample split + reduce) on CPU.

The renderer behind the `Sampler` face is the CPU oracle here (a test may use it); the code
under test is micro_raytracer_b200/distributed.py, the same functions bench.py runs over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from micro_raytracer_b200.distributed import passes_of_rank, render_distributed


def test_passes_of_rank_partitions_every_sample_once():
    for spp in (0, 1, 7, 8, 1024, 1027):
        for world in (1, 2, 3, 8):
            assert sum(passes_of_rank(spp, r, world) for r in range(world)) == spp
    assert [passes_of_rank(10, r, 4) for r in range(4)] == [3, 3, 2, 2]
    with pytest.raises(ValueError):
        passes_of_rank(4, 2, 2)


def _worker(rank, world, port, spp, out_path):
    import oracle_lib
    from util import load
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r = load("CornellBox2", (24, 24), 2.0)
        s = oracle_lib.OracleSampler(workers=1)
        total = render_distributed(s, r.scene, r.frame, r.rt, spp)
        assert (total is not None) == (rank == 0)
        if rank == 0:
            np.save(out_path, total)
    finally:
        dist.destroy_process_group()


def test_two_rank_sample_split_equals_single_render(tmp_path):
    import oracle_lib
    from util import load
    spp = 5  # odd: the ranks render 3 and 2 passes
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(2, port, spp, out), nprocs=2, join=True)
    got = np.load(out)
    r = load("CornellBox2", (24, 24), 2.0)
    s = oracle_lib.OracleSampler(workers=1)
    s.execute(r.scene, r.frame, r.rt, spp)
    want, n = s.accum()
    assert n == spp
    # same paths (the RNG is keyed by the global sample index), different summation order
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
