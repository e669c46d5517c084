"""GPU parity at the sizes BASELINE.json names (run with -m gpu): the CUDA path, through the C ABI, against
the CPU oracle on the benchmark workload itself (256^3, sampled sources), on config 2 (128^3, the ten sources
of inputs/test_sources_standard.dat), on config 1 (300^3, one 1e57 source) and on the 512^3 mesh whose last
layer the reference never traces.

Tolerances (BASELINE.json north_star): ionized fractions abs 1e-6, rates rel 1e-6, photon statistics rel 1e-6;
subbox counts, update counts and the set of cells with a rate exact."""
import os
import sys

import numpy as np
import pytest

from problems import make_problem, setup_oracle, setup_gpu

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu
RATE_RTOL = 1e-6
X_ATOL = 1e-6
YEAR = 3.15576e7

# inputs/test_sources_standard.dat of the reference (recipe 7: column 4 = photons/s, NormFlux = col4/S_star,
# sourceprops.F90:296,379-381): i j k flux
SOURCES_STANDARD = [(50, 50, 50, 1e55), (51, 50, 50, 1e55), (52, 50, 50, 1e55), (53, 50, 50, 1e55),
                    (20, 10, 10, 1e57), (70, 70, 50, 1e55), (72, 70, 50, 1e55), (70, 72, 50, 1e55),
                    (72, 72, 50, 1e56), (20, 10, 90, 1e54)]
# inputs/test_sources_onesrc.dat
SOURCE_ONE = [(50, 50, 50, 1e57)]


def _route(monkeypatch, routing):
    if routing == "cta":
        monkeypatch.setenv("C2B_CLUSTER_MIN_NBOX", "100000")
    elif routing == "cluster":
        monkeypatch.setenv("C2B_CLUSTER_MIN_NBOX", "0")
        monkeypatch.setenv("C2B_CLUSTER_MAX_SOURCES", "1000000")
    else:
        monkeypatch.delenv("C2B_CLUSTER_MIN_NBOX", raising=False)
        monkeypatch.delenv("C2B_CLUSTER_MAX_SOURCES", raising=False)


def _rates_close(gpu, cpu, rtol=RATE_RTOL):
    gpu = np.asarray(gpu).reshape(-1)
    cpu = np.asarray(cpu).reshape(-1)
    nz = cpu != 0
    assert np.array_equal(gpu == 0, cpu == 0), "sets of cells with a rate differ"
    err = float(np.max(np.abs(gpu[nz] - cpu[nz]) / np.abs(cpu[nz]))) if nz.any() else 0.0
    assert err <= rtol, "max relative rate error %.3e" % err
    return err


@pytest.fixture(scope="module")
def bench_workload():
    import bench
    return bench.build_workload(256, 10000, 25.0)


@pytest.mark.parametrize("routing", ["cta", "cluster"])
def test_bench_workload_sampled_sources(bench_workload, routing, monkeypatch):
    """BASELINE configs[2] as bench.py times it (256^3 log-normal density, clumping grid, LLS, the
    mid-reionization bubble state): 32 of the 10^4 sources (the three brightest + every 344th), one
    pass_all_sources + one global_pass, against the oracle -- with each of the two ray-trace kernels"""
    from oracle import oracle as O
    import c2ray3dm_b200 as pkg
    _route(monkeypatch, routing)
    w = bench_workload
    ns = len(w["normflux"])
    sel = np.unique(np.concatenate([[0, 1, 2], np.arange(0, ns, ns // 29)]))[:32]
    assert len(sel) == 32
    dt = 0.5e6 * YEAR
    o = O.Oracle(256)
    o.set_density(w["ndens"])
    o.set_geometry(w["dr"], w["vol"])
    o.set_clumping(5, 1.0, w["clumping"])
    o.set_lls(True, 1, w["coldensh_LLS"], None, 0.0)
    o.set_sources(w["srcpos"][sel], w["normflux"][sel], 1e48)
    o.set_xh(w["xh"])
    o.set_threads(os.cpu_count() or 1)
    o.xh_av[...] = w["xh"]
    o.xh_intermed[...] = w["xh"]
    o.state_before()
    o.set_rates_to_zero()
    ro = o.pass_all_sources()

    e = pkg.Evolve(256, type_of_clumping=5, use_LLS=True, type_of_LLS=1)
    e.rad_ini()
    e.set_density(w["ndens"])
    e.set_geometry(w["dr"], w["vol"])
    e.set_clumping(w["clumping"])
    e.set_LLS(coldensh_LLS=w["coldensh_LLS"])
    e.set_sources(w["srcpos"][sel], w["normflux"][sel])
    e.set_xh(w["xh"])
    e.begin_step()
    rg = e.pass_all_sources()
    assert rg.updates == ro.updates and rg.sum_nbox_all == ro.sum_nbox_all
    assert rg.updates > 32 * 40 ** 3                               # the traces are long ones
    # per-source subbox counts of the three brightest and three of the faintest sampled sources: exact
    nbox = e.source_nbox()
    o1 = O.Oracle(256)
    o1.set_density(w["ndens"])
    o1.set_geometry(w["dr"], w["vol"])
    o1.set_lls(True, 1, w["coldensh_LLS"], None, 0.0)
    o1.set_sources(w["srcpos"][sel], w["normflux"][sel], 1e48)
    o1.set_xh(w["xh"])
    o1.xh_av[...] = w["xh"]
    for k in (1, 2, 3, 30, 31, 32):
        o1.set_rates_to_zero()
        assert nbox[k - 1] == o1.do_source(k).nbox
    assert rg.photon_loss_all == pytest.approx(ro.photon_loss_all, rel=RATE_RTOL)
    _rates_close(e.phih_grid, o.phih)
    go = o.global_pass(dt, ro.photon_loss_all)
    gg = e.global_pass(dt)
    assert gg.conv_flag == go.conv_flag
    np.testing.assert_allclose(e.xh_intermed, o.xh_intermed, rtol=0, atol=X_ATOL)
    np.testing.assert_allclose(e.xh_av, o.xh_av, rtol=0, atol=X_ATOL)
    assert gg.sum_xh_intermed == pytest.approx(go.sum_xh_intermed, rel=1e-9)
    for n in ("h0_after", "h1_after", "totrec", "totcollisions", "total_photon_loss", "totalsrc"):
        assert getattr(gg.stats, n) == pytest.approx(getattr(go.stats, n), rel=1e-6), n
    e.close()


def _nbody_test_problem(N, sources):
    """the nbody_test problem of the reference (nbody_test.F90, density_module.F90:129-147): uniform mean
    density at z=9, 100/h Mpc box, xh = 2e-4, T = 1e4 K, LLS type 1; source positions as in the input file"""
    pos = [[s[0], s[1], s[2]] for s in sources]
    p = make_problem(N, nsrc=len(pos), seed=3, state="neutral", use_LLS=True, dens="uniform", srcpos=pos)
    p["normflux"] = np.array([s[3] / 1e48 for s in sources])
    return p


def _run_history(p, nsteps, dt, threads):
    """evolve3D nsteps times on both sides with the host's cosmo_evol between the steps (C2Ray.F90:367-379)"""
    from c2ray3dm_b200 import synthetic as syn
    o = setup_oracle(p)
    o.set_threads(threads)
    e = setup_gpu(p)
    tables = e.rad_ini()
    o.set_tables(*tables)
    ndens = p["ndens"].copy()
    dr, vol = p["dr"].copy(), p["vol"]
    for step in range(nsteps):
        zf = 1.0 + 2e-3 * (step + 1)
        zf3 = zf * zf * zf
        dr, vol = dr * zf, vol * zf3
        ndens = (ndens.astype(np.float64) / zf3).astype(np.float32)
        o.set_density(ndens)
        o.set_geometry(dr, vol)
        o.set_lls(True, 1, syn.lls_coldens(dr[0], 9.0), None, 0.0)
        e.cosmo_evol(zf)
        e.set_LLS(coldensh_LLS=syn.lls_coldens(dr[0], 9.0))
        ro = o.evolve3D(dt)
        rg = e.evolve3D(step * dt, dt)
        assert (rg.niter, rg.converged, rg.conv_criterion) == (ro.niter, ro.converged, ro.conv_criterion)
        assert list(rg.conv_flag[1:rg.niter + 1]) == list(ro.conv_flag[1:ro.niter + 1])
        assert list(rg.sum_nbox_all[1:rg.niter + 1]) == list(ro.sum_nbox_all[1:ro.niter + 1])
        assert rg.total_updates == ro.total_updates
        np.testing.assert_allclose(e.xh, o.xh, rtol=0, atol=X_ATOL)
        _rates_close(e.phih_grid, o.phih)
        for n in ("photcons", "totrec", "totcollisions", "total_photon_loss"):
            assert getattr(rg.final_stats, n) == pytest.approx(getattr(ro.final_stats, n), rel=1e-6), n
        assert rg.grtotal_ion == pytest.approx(ro.grtotal_ion, rel=1e-6)
    e.close()


@pytest.mark.parametrize("routing", ["auto", "cta"])
def test_config2_ten_sources_128_three_steps(routing, monkeypatch):
    """BASELINE configs[1]: 128^3, the ten sources of test_sources_standard.dat, three evolve3D steps of 1 Myr
    from xh = 2e-4 with cosmo_evol between them"""
    _route(monkeypatch, routing)
    p = _nbody_test_problem(128, SOURCES_STANDARD)
    _run_history(p, 3, 1e6 * YEAR, os.cpu_count() or 1)


def test_config1_single_source_300_one_step():
    """BASELINE configs[0] at its real size: the default compiled mesh 300^3, one 1e57 s^-1 source at
    (50,50,50); with one source conv_criterion is 0, so the 1e-4 test on the sums ends the outer loop
    (55 iterations)"""
    p = _nbody_test_problem(300, SOURCE_ONE)
    _run_history(p, 1, 1e6 * YEAR, os.cpu_count() or 1)


def test_full_box_update_count_512_quirk():
    """N=512: R=255=5*51 stops the walk one pass before the -256 layer is reached: 511^3 updates per
    fully-traced source and exactly that many cells with a rate (evolve_source.F90:128-136).
    A zero loss threshold makes the trace cover the box."""
    N = 512
    import c2ray3dm_b200 as pkg
    from c2ray3dm_b200 import synthetic as syn
    e = pkg.Evolve(N, use_LLS=False, loss_fraction=0.0)
    e.rad_ini()
    e.set_density(np.full(N ** 3, syn.avg_dens(9.0), dtype=np.float32))
    dr, vol = syn.proper_geometry(N, 9.0)
    e.set_geometry(dr, vol)
    src = np.array([[100, 200, 300]], dtype=np.int32)
    e.set_sources(src, [1e9])
    e.set_xh(np.full(N ** 3, 1.0 - 1e-6))
    e.begin_step()
    r = e.pass_all_sources()
    assert r.sum_nbox_all == 51
    assert r.updates == 511 ** 3
    ph = e.phih_grid.reshape(N, N, N)
    assert np.count_nonzero(ph) == 511 ** 3
    # the untraced layer is the one at offset -256 (periodic) along each axis
    for axis, s in enumerate((src[0, 2], src[0, 1], src[0, 0])):      # array axes are (k, j, i)
        layer = (s - 1 - 256) % N
        assert not np.take(ph, layer, axis=axis).any()
    e.close()
