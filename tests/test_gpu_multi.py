"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): two ranks, sources dealt round-robin in the first pass
(master_slave.F90:85) and by predicted trace length and measured rank speed afterwards, ncclAllReduce of phih_grid,
of the packed scalars (evolve.F90:577-616) and of the per-source subbox counts, then the per-cell pass on every
rank.  Compared with the single-process oracle and between ranks."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
DT = 1e6 * 3.15576e7


def _ngpu():
    try:
        from c2ray3dm_b200 import lib
        return lib.load().c2b_device_count()
    except Exception:
        return 0


def _worker(rank, world, uid, q, case):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from problems import make_problem, setup_gpu
    case = dict(case)
    steps = case.pop("steps", 1)
    case_steps = steps
    p = make_problem(**case)
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    e = setup_gpu(p, rank=rank, nranks=world, device=rank)
    e.comm_init(uid)
    for step in range(case_steps):
        rep = e.evolve3D(step * DT, DT)
    q.put((rank, e.xh, e.phih_grid, rep.niter, rep.total_updates, list(rep.conv_flag[1:rep.niter + 1]),
           rep.final_stats.photcons, rep.photon_loss_all[1], e.source_nbox(), e.source_owner()))
    e.close()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("case", [dict(N=32, nsrc=9, seed=17, state="random", use_LLS=True, clumping="grid"),
                                  dict(N=32, nsrc=25, seed=23, state="random", use_LLS=True, flux=3e7, steps=2)],
                         ids=["9src", "25src_2steps"])
def test_two_ranks_match_oracle_and_each_other(case):
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from problems import make_problem, setup_oracle
    from c2ray3dm_b200 import Evolve
    steps = case.pop("steps", 1)
    wcase = dict(case, steps=steps)
    uid = Evolve.get_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, uid, q, wcase)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = make_problem(**case)
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    o = setup_oracle(p)
    for step in range(steps):
        ro = o.evolve3D(DT)
    for rank in (0, 1):
        _, xh, phih, niter, upd, conv, photcons, loss1, nbox, owner = res[rank]
        assert niter == ro.niter and upd == ro.total_updates
        assert conv == list(ro.conv_flag[1:ro.niter + 1])
        np.testing.assert_allclose(xh, o.xh, rtol=0, atol=1e-6)
        nz = o.phih != 0
        assert np.max(np.abs(phih[nz] - o.phih[nz]) / o.phih[nz]) < 1e-6
        assert photcons == pytest.approx(ro.final_stats.photcons, rel=1e-6)
        assert loss1 == pytest.approx(ro.photon_loss_all[1], rel=1e-6)
        # every rank knows the subbox count of every source (all-reduced: the next pass is dealt by them)
        assert int(np.sum(nbox)) == ro.sum_nbox_all[ro.niter]
        assert set(owner.tolist()) == {0, 1}
    np.testing.assert_array_equal(res[0][8], res[1][8])
    np.testing.assert_array_equal(res[0][9], res[1][9])   # both ranks computed the same assignment
    # replicas are bit-identical after the all-reduce (every rank runs the same per-cell pass)
    np.testing.assert_array_equal(res[0][1], res[1][1])
    np.testing.assert_array_equal(res[0][2], res[1][2])
