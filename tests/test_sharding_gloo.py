"""world_size-2 gloo test of the multi-GPU scheme's host logic: sources dealt round-robin to ranks
(master_slave.F90:85), partial rate grids summed by all-reduce (evolve.F90:599-602), the three scalars
packed into one small all-reduce (:587-613), then every rank runs the same per-cell pass."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from problems import make_problem, setup_oracle
    from c2ray3dm_b200 import shard_sources
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = make_problem(16, nsrc=7, seed=31, state="random")
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    o = setup_oracle(p)
    o.set_rank(rank, world)
    o.xh_av[...] = p["xh"]
    o.set_rates_to_zero()
    r = o.pass_all_sources()
    mine = shard_sources(7, rank, world)
    phih = torch.from_numpy(o.phih.copy())
    small = torch.tensor([r.photon_loss_all, float(r.sum_nbox_all), float(r.updates)], dtype=torch.float64)
    dist.all_reduce(phih)
    dist.all_reduce(small)
    o.phih[...] = phih.numpy()
    g = o.global_pass(1e13, small[0].item())
    q.put((rank, mine, phih.numpy(), small.numpy(), o.xh_intermed.copy(), g.conv_flag))
    dist.barrier()
    dist.destroy_process_group()


def test_source_sharding_two_ranks_gloo():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from problems import make_problem, setup_oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert res[0][1] == [1, 3, 5, 7] and res[1][1] == [2, 4, 6]
    # single-rank answer
    p = make_problem(16, nsrc=7, seed=31, state="random")
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    o = setup_oracle(p)
    o.xh_av[...] = p["xh"]
    o.set_rates_to_zero()
    r = o.pass_all_sources()
    g = o.global_pass(1e13, r.photon_loss_all)
    for rank in (0, 1):
        np.testing.assert_allclose(res[rank][2], o.phih, rtol=1e-12, atol=0)
        assert res[rank][3][0] == pytest.approx(r.photon_loss_all, rel=1e-12)
        assert int(res[rank][3][1]) == r.sum_nbox_all and int(res[rank][3][2]) == r.updates
        np.testing.assert_allclose(res[rank][4], o.xh_intermed, rtol=0, atol=1e-12)
        assert res[rank][5] == g.conv_flag
    # replicas are bit-identical after the all-reduce
    np.testing.assert_array_equal(res[0][2], res[1][2])
    np.testing.assert_array_equal(res[0][4], res[1][4])


def _worker_deal(rank, world, port, q):
    """two passes the way the library runs them with nranks > 1: static split, all-reduce of the per-source subbox
    counts, then the sources dealt again by c2b_deal_sources (the product's host rule) from the all-reduced inputs"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import torch
    import torch.distributed as dist
    from problems import make_problem, setup_oracle
    from c2ray3dm_b200 import lib as L
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nsrc = 11
    p = make_problem(16, nsrc=nsrc, seed=33, state="random", flux=3e7)
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    o = setup_oracle(p)
    o.xh_av[...] = p["xh"]
    # pass 1: ns = 1+rank, 1+rank+npr, ... (master_slave.F90:85); every rank fills its own subbox counts
    o.set_rates_to_zero()
    nbox = torch.zeros(nsrc, dtype=torch.int64)
    upd = torch.zeros(nsrc, dtype=torch.int64)
    for ns in range(1 + rank, nsrc + 1, world):
        r = o.do_source(ns)
        nbox[ns - 1], upd[ns - 1] = r.nbox, r.updates
    dist.all_reduce(nbox)
    dist.all_reduce(upd)
    # the same inputs on both ranks -> the same assignment; rank 1 is taken to be half as fast
    cost = np.ascontiguousarray(upd.numpy(), dtype=np.int64)
    speed = np.array([1.0, 0.5])
    owner = np.full(nsrc, -1, dtype=np.int32)
    rc = L.load().c2b_deal_sources(nsrc, cost.ctypes.data_as(C.POINTER(C.c_int64)), world,
                                   speed.ctypes.data_as(C.POINTER(C.c_double)), owner.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    # pass 2 with the dealt sources
    o.set_rates_to_zero()
    for ns in range(1, nsrc + 1):
        if owner[ns - 1] == rank:
            o.do_source(ns)
    phih = torch.from_numpy(o.phih.copy())
    dist.all_reduce(phih)
    q.put((rank, owner.copy(), cost.copy(), phih.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_dealt_sources_two_ranks_gloo():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from problems import make_problem, setup_oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29100 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker_deal, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    np.testing.assert_array_equal(res[0][1], res[1][1])          # both ranks computed the same assignment
    owner, cost = res[0][1], res[0][2]
    assert set(owner.tolist()) == {0, 1}
    share0 = cost[owner == 0].sum() / cost.sum()
    assert abs(share0 - 2.0 / 3.0) < cost.max() / cost.sum()      # shares follow the speeds to within one source
    p = make_problem(16, nsrc=11, seed=33, state="random", flux=3e7)
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    o = setup_oracle(p)
    o.xh_av[...] = p["xh"]
    o.set_rates_to_zero()
    o.pass_all_sources()
    for rank in (0, 1):
        np.testing.assert_allclose(res[rank][3], o.phih, rtol=1e-12, atol=0)
    np.testing.assert_array_equal(res[0][3], res[1][3])
