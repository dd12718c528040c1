"""Random-scene checks that need no GPU: the loader/packer accept every scene the generator makes,
and the oracle's two evaluation orders of reduce_light (as written vs forward) agree on them."""
import numpy as np
import pytest

import micro_raytracer_b200 as mrt
import oracle_lib
from fuzz_scenes import random_scene


@pytest.mark.parametrize("seed", range(12))
def test_oracle_literal_equals_forward_on_random_scenes(seed):
    r = random_scene(seed, res=(28, 20))
    packed = mrt.pack_scene(r.scene)
    assert packed.c.n_instances >= 2
    a = oracle_lib.OracleSampler(mode=oracle_lib.LITERAL)
    b = oracle_lib.OracleSampler(mode=oracle_lib.FORWARD)
    for s in (a, b):
        s.execute(r.scene, r.frame, r.rt, 2)
    ia, ib = a.accum()[0], b.accum()[0]
    fin = np.isfinite(ia).all(axis=2) & np.isfinite(ib).all(axis=2)
    assert fin.mean() > 0.98
    np.testing.assert_allclose(ia[fin], ib[fin], rtol=5e-5, atol=5e-6)
