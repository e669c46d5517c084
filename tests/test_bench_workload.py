"""CPU tests of bench.py's weak-scaling construction: N periodic copies of the base volume in one mesh give every
source exactly the environment it has in the base volume, so the rate grid of the tiled job is the tiled rate grid of
the base job and every rank has the 1-GPU updates to do."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import oracle as O  # noqa: E402


def test_tiling_factors():
    assert [bench.tiling(n) for n in (1, 2, 4, 8)] == [(1, 1, 1), (1, 1, 2), (1, 2, 2), (2, 2, 2)]
    for n in range(1, 17):
        t = bench.tiling(n)
        assert t[0] * t[1] * t[2] == n


def _pass(w, mesh):
    o = O.Oracle(mesh)
    o.set_density(w["ndens"])
    o.set_geometry(w["dr"], w["vol"])
    o.set_clumping(5, 1.0, w["clumping"])
    o.set_lls(True, 1, w["coldensh_LLS"], None, 0.0)
    o.set_sources(w["srcpos"], w["normflux"], 1e48)
    o.set_xh(w["xh"])
    o.xh_av[...] = w["xh"]
    o.xh_intermed[...] = w["xh"]
    o.state_before()
    o.set_rates_to_zero()
    r = o.pass_all_sources()
    return r, o.phih.copy()


def test_tiled_job_is_copies_of_the_base_job():
    base, nsrc, bubble = 32, 6, 3.0
    w1 = bench._build_workload(base, nsrc, bubble)
    r1, ph1 = _pass(w1, base)
    # the traces must end inside the base half box for the copies to be independent of the mesh size
    assert r1.updates < nsrc * (base - 1) ** 3
    for world in (2, 4):
        t = bench.tiling(world)
        mesh = (base * t[0], base * t[1], base * t[2])
        w = bench._build_workload(mesh, nsrc * world, bubble, t)
        assert w["ndens"].shape == (mesh[2], mesh[1], mesh[0]) and len(w["normflux"]) == nsrc * world
        # copies of a source are consecutive: the static round-robin gives every rank one copy of every source
        assert np.array_equal(w["normflux"].reshape(nsrc, world), np.repeat(w1["normflux"], world).reshape(nsrc, world))
        r, ph = _pass(w, mesh)
        assert r.updates == world * r1.updates and r.sum_nbox_all == world * r1.sum_nbox_all
        np.testing.assert_allclose(ph, np.tile(ph1, (t[2], t[1], t[0])), rtol=1e-12, atol=0)
