"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against
the CPU oracle on the same seeded inputs, against the committed golden fixtures, and -- at sizes the
oracle cannot reach -- through size-independent properties.

Tolerances (BASELINE.json north_star): ionized fractions abs 1e-6, photo-ionization rates rel 1e-6,
photon statistics rel 1e-6; shell/cell indexing, nbox and update counts exact."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from problems import make_problem, setup_oracle, setup_gpu

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RATE_RTOL = 1e-6
X_ATOL = 1e-6
DT = 1e6 * 3.15576e7


@pytest.fixture(params=["auto", "cta", "cluster", "warp"], autouse=True)
def routing(request, monkeypatch):
    """every test runs with the default routing of sources to the ray-trace kernels, with the
    one-CTA-per-source kernel only, with the cluster-of-6 kernel only, and with the one-warp-per-source kernel
    taking every source whose previous trace ended after one subbox (the rest: one CTA per source)"""
    monkeypatch.delenv("C2B_WARP_MIN_SOURCES", raising=False)
    # the y-fastest twin grids of the x-principal faces: forced on with "cta", off with "cluster", by size otherwise
    monkeypatch.delenv("C2B_TWINS", raising=False)
    if request.param == "cta":
        monkeypatch.setenv("C2B_TWINS", "1")
    elif request.param == "cluster":
        monkeypatch.setenv("C2B_TWINS", "0")
    if request.param == "warp":
        monkeypatch.setenv("C2B_CLUSTER_MIN_NBOX", "100000")
        monkeypatch.setenv("C2B_DEBUG_CLUSTER", "0")
        monkeypatch.setenv("C2B_WARP_MIN_SOURCES", "1")
    elif request.param == "cta":
        monkeypatch.setenv("C2B_CLUSTER_MIN_NBOX", "100000")
        monkeypatch.setenv("C2B_DEBUG_CLUSTER", "0")
    elif request.param == "cluster":
        monkeypatch.setenv("C2B_CLUSTER_MIN_NBOX", "0")
        monkeypatch.setenv("C2B_CLUSTER_MAX_SOURCES", "1000000")
        monkeypatch.setenv("C2B_DEBUG_CLUSTER", "1")
    else:
        monkeypatch.delenv("C2B_CLUSTER_MIN_NBOX", raising=False)
        monkeypatch.delenv("C2B_CLUSTER_MAX_SOURCES", raising=False)
        monkeypatch.delenv("C2B_DEBUG_CLUSTER", raising=False)
    return request.param


def _rates_close(gpu, cpu, rtol=RATE_RTOL):
    """relative tolerance on every cell that has a rate; cells the oracle leaves at exactly 0 must be 0"""
    gpu = np.asarray(gpu).reshape(-1)
    cpu = np.asarray(cpu).reshape(-1)
    nz = cpu != 0
    assert np.array_equal(gpu == 0, cpu == 0), "sets of cells with a rate differ"
    err = np.max(np.abs(gpu[nz] - cpu[nz]) / np.abs(cpu[nz])) if nz.any() else 0.0
    assert err <= rtol, "max relative rate error %.3e" % err
    return err


CASES = [
    dict(N=16, nsrc=1, seed=1, state="ionized", use_LLS=False),
    dict(N=21, nsrc=3, seed=5, state="ionized", use_LLS=False),               # odd mesh
    dict(N=24, nsrc=3, seed=5, state="ionized", use_LLS=True),
    dict(N=(16, 20, 12), nsrc=3, seed=5, state="ionized", use_LLS=False),      # non-cubic
    dict(N=22, nsrc=2, seed=6, state="ionized", use_LLS=False),               # the -N/2 layer quirk
    dict(N=32, nsrc=8, seed=7, state="random", use_LLS=True),
    dict(N=32, nsrc=8, seed=7, state="random", use_LLS=True, type_of_LLS=2, clumping="grid"),
    dict(N=32, nsrc=8, seed=7, state="ionized", use_LLS=True, type_of_LLS=3),
    dict(N=32, nsrc=5, seed=8, state="neutral", use_LLS=True),                 # every ray stops in subbox 1
    dict(N=48, nsrc=64, seed=9, state="random", use_LLS=True, clumping="scalar2"),
]


def _problem(c):
    p = make_problem(**c)
    if c["state"] == "random":
        p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    return p


@pytest.fixture(scope="module")
def gpu_tables():
    p = make_problem(8)
    e = setup_gpu(p)
    t = e.rad_ini()
    e.close()
    return t


def test_tables_match_oracle(gpu_tables):
    """rad_ini on the device vs the restated rad_ini (exp/pow differ by an ulp at most)"""
    from oracle import oracle as O
    thick, thin, _ = O.rad_ini()
    np.testing.assert_allclose(gpu_tables[0], thick, rtol=1e-12, atol=0)
    np.testing.assert_allclose(gpu_tables[1], thin, rtol=1e-12, atol=0)
    g = np.load(os.path.join(GOLD, "tables.npz"))
    np.testing.assert_allclose(gpu_tables[0][g["idx"]], g["thick"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "N%s_%s_s%d" % (c["N"], c["state"], c["nsrc"]))
def test_raytrace_pass_matches_oracle(case, gpu_tables):
    p = _problem(case)
    e = setup_gpu(p, tables=gpu_tables)
    o = setup_oracle(p, tables=gpu_tables)
    o.xh_av[...] = p["xh"]
    o.set_rates_to_zero()
    r = o.pass_all_sources()
    e.begin_step()
    g = e.pass_all_sources()
    assert g.sum_nbox_all == r.sum_nbox_all           # shell / subbox indexing: exact
    assert g.updates == r.updates                     # gate count: exact
    assert g.photon_loss_all == pytest.approx(r.photon_loss_all, rel=RATE_RTOL)
    _rates_close(e.phih_grid, o.phih)
    # per source: subbox count exact, boundary loss within tolerance
    nbox = e.source_nbox()
    loss = e.source_loss()
    for ns in range(1, len(p["normflux"]) + 1):
        o1 = setup_oracle(p, tables=gpu_tables)
        o1.xh_av[...] = p["xh"]
        o1.set_rates_to_zero()
        rr = o1.do_source(ns)
        assert nbox[ns - 1] == rr.nbox
        assert loss[ns - 1] == pytest.approx(rr.photon_loss_src, rel=RATE_RTOL, abs=0)
        if ns > 3:
            break
    e.close()


@pytest.mark.parametrize("case", CASES[:7], ids=lambda c: "N%s_%s" % (c["N"], c["state"]))
def test_single_source_column_densities(case, gpu_tables):
    """do_source for one source: the whole coldensh_out grid (rel 1e-12), the set of traced cells
    (exact), the rates, nbox and the boundary loss"""
    p = _problem(case)
    e = setup_gpu(p, tables=gpu_tables)
    e.begin_step()
    o = setup_oracle(p, tables=gpu_tables)
    o.xh_av[...] = p["xh"]
    o.set_rates_to_zero()
    rr = o.do_source(1)
    cd, ph, nbox, loss = e.trace_source_debug(1)
    assert nbox == rr.nbox
    assert np.array_equal(cd == 0, o.coldensh_out == 0)
    np.testing.assert_allclose(cd, o.coldensh_out, rtol=1e-12, atol=0)
    _rates_close(ph, o.phih)
    assert loss == pytest.approx(rr.photon_loss_src, rel=RATE_RTOL)
    e.close()


@pytest.mark.parametrize("case", [CASES[0], CASES[2], CASES[5], CASES[6], CASES[9]],
                         ids=lambda c: "N%s_%s" % (c["N"], c["state"]))
def test_global_pass_matches_oracle(case, gpu_tables):
    """the per-cell kernel alone: both start from the oracle's rate grid"""
    p = _problem(case)
    o = setup_oracle(p, tables=gpu_tables)
    o.state_before()
    o.xh_av[...] = p["xh"]
    o.xh_intermed[...] = p["xh"]
    o.set_rates_to_zero()
    r = o.pass_all_sources()
    e = setup_gpu(p, tables=gpu_tables)
    e.begin_step()
    e.set_iter_state(1, r.photon_loss_all, o.phih, o.xh_av, o.xh_intermed)
    gg = e.global_pass(DT)
    go = o.global_pass(DT, r.photon_loss_all)
    assert gg.conv_flag == go.conv_flag
    assert gg.min_avg_neutral == pytest.approx(go.min_avg_neutral, rel=1e-12)
    np.testing.assert_allclose(e.xh_intermed, o.xh_intermed, rtol=0, atol=1e-12)
    np.testing.assert_allclose(e.xh_av, o.xh_av, rtol=0, atol=1e-12)
    assert gg.sum_xh_intermed == pytest.approx(go.sum_xh_intermed, rel=1e-12)
    for n in ("h0_after", "h1_after", "totrec", "totcollisions", "total_photon_loss", "totalsrc"):
        assert getattr(gg.stats, n) == pytest.approx(getattr(go.stats, n), rel=1e-11), n
    e.close()


@pytest.mark.parametrize("case", CASES, ids=lambda c: "N%s_%s_s%d" % (c["N"], c["state"], c["nsrc"]))
def test_evolve3d_step_matches_oracle(case, gpu_tables):
    """one full evolve3D call: iteration count, per-iteration convergence counters, final fractions,
    rates and the photon-conservation statistics"""
    p = _problem(case)
    o = setup_oracle(p, tables=gpu_tables)
    ro = o.evolve3D(DT)
    e = setup_gpu(p, tables=gpu_tables)
    rg = e.evolve3D(0.0, DT)
    assert (rg.niter, rg.converged, rg.conv_criterion) == (ro.niter, ro.converged, ro.conv_criterion)
    assert list(rg.conv_flag[1:rg.niter + 1]) == list(ro.conv_flag[1:ro.niter + 1])
    assert list(rg.sum_nbox_all[1:rg.niter + 1]) == list(ro.sum_nbox_all[1:ro.niter + 1])
    assert rg.total_updates == ro.total_updates
    assert rg.kernel_launches >= 3 * rg.niter
    np.testing.assert_allclose(e.xh, o.xh, rtol=0, atol=X_ATOL)
    np.testing.assert_allclose(e.xh_av, o.xh_av, rtol=0, atol=X_ATOL)
    np.testing.assert_allclose(e.xh_intermed, o.xh_intermed, rtol=0, atol=X_ATOL)
    _rates_close(e.phih_grid, o.phih)
    for n in ("totrec", "totcollisions", "dh0", "total_ion", "photcons", "total_photon_loss", "totalsrc",
              "h0_before", "h1_after"):
        a, b = getattr(rg.final_stats, n), getattr(ro.final_stats, n)
        assert a == pytest.approx(b, rel=1e-6, abs=1e-6 * abs(ro.final_stats.totrec) if n in ("dh0", "total_ion") else 0), n
    assert rg.grtotal_src == pytest.approx(ro.grtotal_src, rel=1e-12)
    e.close()


def test_history_three_steps_with_cosmology(gpu_tables):
    """three consecutive steps with the host's cosmo_evol between them (C2Ray.F90:367-379)"""
    p = _problem(dict(N=24, nsrc=6, seed=14, state="neutral", use_LLS=True, flux=3e8))
    o = setup_oracle(p, tables=gpu_tables)
    e = setup_gpu(p, tables=gpu_tables)
    ndens = p["ndens"].copy()
    dr, vol = p["dr"].copy(), p["vol"]
    for step in range(3):
        zf = 1.0 + 0.01 * (step + 1)
        zf3 = zf * zf * zf
        dr = dr * zf
        vol = vol * zf3
        ndens = (ndens.astype(np.float64) / zf3).astype(np.float32)
        o.set_density(ndens)
        o.set_geometry(dr, vol)
        e.cosmo_evol(zf)
        ro = o.evolve3D(DT)
        rg = e.evolve3D(step * DT, DT)
        assert rg.niter == ro.niter
        np.testing.assert_allclose(e.xh, o.xh, rtol=0, atol=X_ATOL)
        assert rg.final_stats.photcons == pytest.approx(ro.final_stats.photcons, rel=1e-6)
        assert rg.grtotal_ion == pytest.approx(ro.grtotal_ion, rel=1e-6)
    e.close()


@pytest.mark.parametrize("name", ["g16_lls", "g20_clump", "g12x16x10"])
def test_against_committed_golden(name):
    """the CUDA path against the committed fixtures (generated by tests/golden/make_golden.py)"""
    sys.path.insert(0, GOLD)
    import make_golden as mg
    g = np.load(os.path.join(GOLD, name + ".npz"))
    p = mg.make_case(name)
    e = setup_gpu(p)
    e.begin_step()
    r = e.pass_all_sources()
    assert r.sum_nbox_all == g["sum_nbox"] and r.updates == g["updates"]
    _rates_close(e.phih_grid, g["phih_pass"])
    assert r.photon_loss_all == pytest.approx(float(g["photon_loss_all"]), rel=RATE_RTOL)
    e.set_xh(p["xh"])
    rep = e.evolve3D(0.0, mg.DT)
    assert rep.niter == g["niter"]
    assert list(rep.conv_flag[:rep.niter + 1])[1:] == list(g["conv_flag"])[1:]
    np.testing.assert_allclose(e.xh, g["xh"], rtol=0, atol=X_ATOL)
    _rates_close(e.phih_grid, g["phih"])
    assert rep.final_stats.photcons == pytest.approx(float(g["photcons"]), rel=1e-6)
    e.close()


def test_restart_from_iteration_dump(gpu_tables):
    """write_iteration_dump / start_from_dump (evolve.F90:285-426).  The reference dumps between pass_all_sources
    and global_pass; the shim does the same with the fine-grained calls.  The dump record taken after the pass of
    iteration 2 matches the oracle's, and a fresh handle restarted from it (restart /= 0: global_pass, then the
    loop continues at niter+1) finishes exactly like the uninterrupted run, and like the oracle's restart."""
    p = _problem(CASES[5])
    # uninterrupted runs
    e = setup_gpu(p, tables=gpu_tables)
    full = e.evolve3D(0.0, DT)
    x_full, xav_full, ph_full = e.xh, e.xh_av, e.phih_grid
    o = setup_oracle(p, tables=gpu_tables)
    o.set_dump_iteration(2)
    ro = o.evolve3D(DT)
    assert full.niter == ro.niter > 2
    dump_o = o.get_dump()
    # the host's loop up to the dump point of iteration 2 (fortran/evolve_b200.F90)
    e2 = setup_gpu(p, tables=gpu_tables)
    e2.begin_step()
    for niter in (1, 2):
        e2.pass_all_sources(niter, DT)
        if niter == 2:
            dump_g = e2.get_iter_state()
        e2.global_pass(DT)
    assert dump_g[0] == dump_o[0] == 2
    assert dump_g[1] == pytest.approx(dump_o[1], rel=RATE_RTOL)
    _rates_close(dump_g[2], dump_o[2])
    np.testing.assert_allclose(dump_g[3], dump_o[3], rtol=0, atol=X_ATOL)
    np.testing.assert_allclose(dump_g[4], dump_o[4], rtol=0, atol=X_ATOL)
    # restart a fresh handle (xh = start of the step, as the xfrac file holds it) from the GPU's dump
    e3 = setup_gpu(p, tables=gpu_tables)
    e3.set_iter_state(*dump_g)
    rest = e3.evolve3D(0.0, DT, restart=1)
    assert (rest.niter, rest.converged) == (full.niter, full.converged)
    assert list(rest.conv_flag[3:rest.niter + 1]) == list(full.conv_flag[3:full.niter + 1])
    np.testing.assert_allclose(e3.xh, x_full, rtol=0, atol=1e-12)
    np.testing.assert_allclose(e3.xh_av, xav_full, rtol=0, atol=1e-12)
    # the last rate grid was accumulated with atomics in both runs: equal up to the order of the additions
    np.testing.assert_allclose(e3.phih_grid, ph_full, rtol=1e-10, atol=0)
    assert rest.final_stats.photcons == pytest.approx(full.final_stats.photcons, rel=1e-10)
    # the oracle restarted from ITS dump agrees too
    o2 = setup_oracle(p, tables=gpu_tables)
    r2 = o2.evolve3D_restart(DT, *dump_o)
    assert (rest.niter, rest.converged) == (r2.niter, r2.converged)
    assert list(rest.conv_flag[3:rest.niter + 1]) == list(r2.conv_flag[3:r2.niter + 1])
    np.testing.assert_allclose(e3.xh, o2.xh, rtol=0, atol=X_ATOL)
    assert rest.final_stats.photcons == pytest.approx(r2.final_stats.photcons, rel=1e-6)
    # the dump itself round-trips exactly through the ABI
    n3, pl3, ph3, xa3, xi3 = (e3.set_iter_state(*dump_g), e3.get_iter_state())[1]
    assert (n3, pl3) == (dump_g[0], dump_g[1])
    for u, v in ((dump_g[2], ph3), (dump_g[3], xa3), (dump_g[4], xi3)):
        np.testing.assert_array_equal(u, v)
    for x in (e, e2, e3):
        x.close()


def test_abi_error_paths_on_device():
    from c2ray3dm_b200 import lib as L, C2RayError
    p = make_problem(12, nsrc=2, seed=2)
    e = setup_gpu(p)
    lib = e.L
    assert lib.c2b_set_density(e.h, None) != 0 and b"null" in lib.c2b_last_error(e.h)
    assert lib.c2b_set_tables(e.h, None, None, 2001) != 0
    t = np.zeros(10)
    assert lib.c2b_set_tables(e.h, t.ctypes.data_as(C.POINTER(C.c_double)), t.ctypes.data_as(C.POINTER(C.c_double)), 10) != 0
    bad = np.array([[0, 1, 1]], dtype=np.int32)
    with pytest.raises(C2RayError):
        e.set_sources(bad, [1.0])
    with pytest.raises(C2RayError):
        e.set_sources(np.array([[13, 1, 1]], dtype=np.int32), [1.0])
    rep = L.StepReport()
    assert lib.c2b_evolve3d(e.h, 0.0, -1.0, 0, C.byref(rep)) != 0
    assert lib.c2b_evolve3d(e.h, 0.0, 1.0, 1, C.byref(rep)) != 0      # restart without state
    # handle still usable afterwards
    e.set_sources(p["srcpos"], p["normflux"])
    assert e.evolve3D(0.0, DT).niter >= 1
    # missing inputs are reported, not crashed on
    e2 = __import__("c2ray3dm_b200").Evolve(12)
    with pytest.raises(C2RayError) as ei:
        e2.evolve3D(0.0, DT)
    assert "not set" in str(ei.value)
    e.close()
    e2.close()


def test_zero_and_empty_sources(gpu_tables):
    p = _problem(dict(N=16, nsrc=3, seed=4, state="ionized"))
    p["normflux"][1] = 0.0                      # never traced: adds 0 to sum_nbox
    e = setup_gpu(p, tables=gpu_tables)
    o = setup_oracle(p, tables=gpu_tables)
    o.xh_av[...] = p["xh"]
    o.set_rates_to_zero()
    r = o.pass_all_sources()
    e.begin_step()
    g = e.pass_all_sources()
    assert g.sum_nbox_all == r.sum_nbox_all and e.source_nbox()[1] == 0
    _rates_close(e.phih_grid, o.phih)
    e.set_sources(np.zeros((0, 3), dtype=np.int32), np.zeros(0))
    g = e.pass_all_sources()
    assert g.updates == 0 and not e.phih_grid.any()
    e.close()


def test_phih_single_precision_output(gpu_tables):
    p = _problem(CASES[2])
    e = setup_gpu(p, tables=gpu_tables)
    e.begin_step()
    e.pass_all_sources()
    np.testing.assert_array_equal(e.phih_grid_si, e.phih_grid.astype(np.float32))   # real(phih_grid,si)
    e.close()


# ---- properties at sizes the oracle cannot reach ---------------------------------------------------
@pytest.mark.parametrize("N,nsrc,bubble", [(128, 40, 9.0), (256, 60, 14.0)], ids=["128", "256_bench_mesh"])
def test_linearity_in_sources(N, nsrc, bubble, gpu_tables):
    """rates are additive over sources at fixed xh_av (evolve_point.F90:283): trace A, B and A+B on the
    benchmark's kind of inputs (log-normal density, sources at the peaks, bubble state, LLS), including the
    benchmark's mesh size, where the oracle is out of reach"""
    from c2ray3dm_b200 import synthetic as syn
    nd = syn.lognormal_density(N, 9.0, 77)
    pos, nf = syn.sources_at_density_peaks(nd, nsrc, 3e8)
    xh = syn.bubble_state(nd.shape, pos, bubble)
    dr, vol = syn.proper_geometry(N, 9.0)
    half = nsrc // 2
    out = []
    for sel in (slice(0, half), slice(half, nsrc), slice(0, nsrc)):
        e = __import__("c2ray3dm_b200").Evolve(N, use_LLS=True)
        e.set_tables(*gpu_tables)
        e.set_density(nd)
        e.set_geometry(dr, vol)
        e.set_LLS(coldensh_LLS=syn.lls_coldens(dr[0], 9.0))
        e.set_sources(pos[sel], nf[sel])
        e.set_xh(xh)
        e.begin_step()
        r = e.pass_all_sources()
        out.append((e.phih_grid, r, e.source_nbox()))
        e.close()
    np.testing.assert_allclose(out[0][0] + out[1][0], out[2][0], rtol=1e-10, atol=1e-30)
    assert out[0][1].updates + out[1][1].updates == out[2][1].updates
    assert out[0][1].photon_loss_all + out[1][1].photon_loss_all == pytest.approx(out[2][1].photon_loss_all, rel=1e-10)
    assert list(out[0][2]) + list(out[1][2]) == list(out[2][2])       # per-source subbox counts are independent
    assert out[2][1].updates > 4 * nsrc * 1331                         # the traces do leave the first subbox


def test_translation_invariance_periodic(gpu_tables):
    """periodic mesh: shifting density, state and sources by the same lattice vector shifts the rates"""
    from c2ray3dm_b200 import synthetic as syn
    N = 64
    nd = syn.lognormal_density(N, 9.0, 5)
    pos, nf = syn.sources_at_density_peaks(nd, 12, 1e8)
    xh = syn.bubble_state(nd.shape, pos, 7.0)
    dr, vol = syn.proper_geometry(N, 9.0)
    shift = (17, 40, 3)  # (i,j,k)

    def run(nd_, xh_, pos_):
        e = __import__("c2ray3dm_b200").Evolve(N, use_LLS=False)
        e.set_tables(*gpu_tables)
        e.set_density(nd_)
        e.set_geometry(dr, vol)
        e.set_sources(pos_, nf)
        e.set_xh(xh_)
        rep = e.evolve3D(0.0, DT)
        res = (e.phih_grid, e.xh, rep.niter)
        e.close()
        return res

    a = run(nd, xh, pos)
    roll = lambda g: np.roll(g, (shift[2], shift[1], shift[0]), axis=(0, 1, 2))
    pos2 = ((pos - 1 + np.array(shift)) % N + 1).astype(np.int32)
    b = run(roll(nd), roll(xh), pos2)
    assert a[2] == b[2]
    np.testing.assert_allclose(roll(a[0]), b[0], rtol=1e-9, atol=1e-30)   # xc=alam*di+real(i0) rounds with i0
    np.testing.assert_allclose(roll(a[1]), b[1], rtol=0, atol=1e-9)


def test_nbody_test_history_single_source(gpu_tables):
    """BASELINE config 1 in miniature: the nbody_test problem (uniform mean density at z=9, 100/h Mpc box,
    LLS type 1, xh=2e-4, T=1e4 K, one 1e57 s^-1 source), 4 consecutive steps with cosmo_evol between them;
    with one source conv_criterion is 0, so only the 1e-4 test on the sums ends the outer loop"""
    from c2ray3dm_b200 import synthetic as syn, constants as K
    N = 40
    p = make_problem(N, nsrc=1, seed=3, state="neutral", use_LLS=True, dens="uniform", srcpos=[[17, 20, 23]])
    p["normflux"] = np.array([1e57 / 1e48])
    o = setup_oracle(p, tables=gpu_tables)
    e = setup_gpu(p, tables=gpu_tables)
    ndens = p["ndens"].copy()
    dr, vol = p["dr"].copy(), p["vol"]
    dt = 1e6 * 3.15576e7
    for step in range(4):
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
        assert ro.conv_criterion == 0 and rg.conv_criterion == 0
        assert rg.niter == ro.niter and rg.converged == ro.converged
        assert list(rg.sum_nbox_all[1:rg.niter + 1]) == list(ro.sum_nbox_all[1:ro.niter + 1])
        np.testing.assert_allclose(e.xh, o.xh, rtol=0, atol=X_ATOL)
        _rates_close(e.phih_grid, o.phih)
        assert rg.final_stats.photcons == pytest.approx(ro.final_stats.photcons, rel=1e-6)
    e.close()


def test_device_pointers_and_probes():
    """harness entry points: device pointers by name, DFMA probe, synchronize"""
    p = make_problem(12, nsrc=1, seed=2)
    e = setup_gpu(p)
    for name in (b"ndens", b"xh", b"xh_av", b"xh_intermed", b"phih"):
        assert e.L.c2b_dev_ptr(e.h, name)
    assert not e.L.c2b_dev_ptr(e.h, b"nonsense")
    e.synchronize()
    rate = e.measure_dfma_rate()
    assert 5e12 < rate < 5e13      # B200: ~1.8e13 FP64 FMA instructions/s
    e.close()


def test_full_box_trace_spills_planes_to_global(gpu_tables):
    """N=100, one source, ionized gas, no LLS: the trace covers the whole periodic box (10 subboxes,
    r up to 50), so the shell planes outgrow shared memory in both ray-trace kernels and live in the
    global scratch; compared cell by cell with the oracle"""
    p = make_problem(100, nsrc=1, seed=23, state="ionized", use_LLS=False, srcpos=[[97, 3, 50]], flux=1e9)
    e = setup_gpu(p, tables=gpu_tables)
    e.begin_step()
    o = setup_oracle(p, tables=gpu_tables)
    o.xh_av[...] = p["xh"]
    o.set_rates_to_zero()
    rr = o.do_source(1)
    cd, ph, nbox, loss = e.trace_source_debug(1)
    assert nbox == rr.nbox == 10
    assert rr.updates == 100 ** 3
    assert np.array_equal(cd == 0, o.coldensh_out == 0)
    np.testing.assert_allclose(cd, o.coldensh_out, rtol=1e-11, atol=0)
    _rates_close(ph, o.phih)
    assert loss == pytest.approx(rr.photon_loss_src, rel=RATE_RTOL)
    g = e.pass_all_sources()
    assert g.updates == 100 ** 3 and g.sum_nbox_all == 10
    _rates_close(e.phih_grid, o.phih)
    e.close()


def test_cluster_kernel_full_step(gpu_tables, monkeypatch):
    """the cluster kernel (six CTAs, one per cube face, boundary loss summed over DSMEM) on every source of a full
    evolve3D step, and its diagnostic trace, against the oracle"""
    monkeypatch.setenv("C2B_CLUSTER_MIN_NBOX", "0")
    monkeypatch.setenv("C2B_CLUSTER_MAX_SOURCES", "1000000")
    monkeypatch.setenv("C2B_DEBUG_CLUSTER", "1")
    p = _problem(dict(N=40, nsrc=5, seed=31, state="random", use_LLS=True))
    e = setup_gpu(p, tables=gpu_tables)
    o = setup_oracle(p, tables=gpu_tables)
    ro = o.evolve3D(DT)
    rg = e.evolve3D(0.0, DT)
    assert rg.niter == ro.niter and rg.total_updates == ro.total_updates
    np.testing.assert_allclose(e.xh, o.xh, rtol=0, atol=X_ATOL)
    _rates_close(e.phih_grid, o.phih)
    cd, ph, nbox, loss = e.trace_source_debug(2)
    o2 = setup_oracle(p, tables=gpu_tables)
    o2.xh_av[...] = e.xh_av
    o2.set_rates_to_zero()
    rr = o2.do_source(2)
    assert nbox == rr.nbox
    np.testing.assert_allclose(cd, o2.coldensh_out, rtol=1e-11, atol=0)
    e.close()


def test_non_cubic_cells(gpu_tables):
    """dr(1) != dr(2) != dr(3): dist2 uses the three cell sizes, the path length only dr(1)
    (evolve_point.F90:166-177); the kernel takes its general-dist2 branch"""
    p = _problem(dict(N=(20, 24, 16), nsrc=4, seed=51, state="random", use_LLS=True))
    p["dr"] = p["dr"] * np.array([1.0, 1.3, 0.8])
    p["vol"] = float(np.prod(p["dr"]))
    e = setup_gpu(p, tables=gpu_tables)
    o = setup_oracle(p, tables=gpu_tables)
    ro = o.evolve3D(DT)
    rg = e.evolve3D(0.0, DT)
    assert rg.niter == ro.niter and rg.total_updates == ro.total_updates
    np.testing.assert_allclose(e.xh, o.xh, rtol=0, atol=X_ATOL)
    _rates_close(e.phih_grid, o.phih)
    assert rg.final_stats.photcons == pytest.approx(ro.final_stats.photcons, rel=1e-6)
    e.close()


@pytest.mark.parametrize("lls,clump", [(1, "scalar"), (2, "grid"), (3, "scalar")], ids=["lls1", "lls_grid", "lls_rmax"])
def test_warp_per_source_kernel_and_hand_over(lls, clump, gpu_tables, monkeypatch, routing):
    """early reionization: many sources whose traces end after one subbox go to the one-warp-per-source kernel; the
    bright ones outgrow the first subbox and are handed over to the one-CTA kernel on the device.  Three steps
    against the oracle; the route counters prove both paths ran."""
    if routing != "warp":
        pytest.skip("runs once, with the warp routing")
    rng = np.random.default_rng(77)
    N, nsrc = 40, 120
    p = make_problem(N, nsrc=nsrc, seed=31, state="neutral", use_LLS=True, type_of_LLS=lls, clumping=clump, flux=2e6)
    p["normflux"][::7] *= 3000.0       # a few bright sources leave the first subbox within these steps
    p["normflux"][5] = 0.0             # a dark source is never traced (evolve_source.F90:128)
    o = setup_oracle(p, tables=gpu_tables)
    e = setup_gpu(p, tables=gpu_tables)
    for step in range(3):
        ro = o.evolve3D(DT)
        rg = e.evolve3D(step * DT, DT)
        assert (rg.niter, rg.converged) == (ro.niter, ro.converged)
        assert list(rg.sum_nbox_all[1:rg.niter + 1]) == list(ro.sum_nbox_all[1:ro.niter + 1])
        assert list(rg.conv_flag[1:rg.niter + 1]) == list(ro.conv_flag[1:ro.niter + 1])
        assert rg.total_updates == ro.total_updates
        np.testing.assert_allclose(e.xh, o.xh, rtol=0, atol=X_ATOL)
        _rates_close(e.phih_grid, o.phih)
        assert rg.final_stats.photcons == pytest.approx(ro.final_stats.photcons, rel=1e-6)
        assert rg.final_stats.total_photon_loss == pytest.approx(ro.final_stats.total_photon_loss, rel=1e-6)
    cta, cluster, warp, handed = e.route_counts()
    assert warp > 0 and handed > 0 and cluster == 0, (cta, cluster, warp, handed)
    nb = e.source_nbox()
    assert nb[5] == 0 and nb.max() >= 2
    e.close()


@pytest.mark.parametrize("subboxsize", [2, 3, 8])
def test_warp_kernel_other_subbox_sizes(subboxsize, gpu_tables, monkeypatch, routing):
    """the oracle carries the reference's subboxsize = 5 (c2ray_parameters.f90:54); for other sizes the
    one-warp-per-source kernel (planes sized from subboxsize, hand-over after the first subbox) must reproduce the
    one-CTA kernel: subbox counts and update counts exactly, rates to rounding (the additions are in another order)"""
    if routing != "warp":
        pytest.skip("runs once, with the warp routing")
    p = make_problem(36, nsrc=60, seed=41, state="neutral", use_LLS=True, flux=2e6)
    p["normflux"][::5] *= 3000.0
    res = []
    for use_warp in (True, False):
        monkeypatch.setenv("C2B_NO_WARP_KERNEL", "0" if use_warp else "1")
        e = setup_gpu(p, tables=gpu_tables, subboxsize=subboxsize)
        for step in range(2):
            rep = e.evolve3D(step * DT, DT)
        res.append((rep.niter, rep.total_updates, list(rep.sum_nbox_all[1:rep.niter + 1]), e.source_nbox().copy(),
                    e.phih_grid.copy(), e.xh.copy(), e.route_counts()))
        e.close()
    (n1, u1, s1, nb1, ph1, x1, rc1), (n2, u2, s2, nb2, ph2, x2, rc2) = res
    assert rc1[2] > 0 and rc2[2] == 0                       # the warp kernel ran in the first run only
    assert (n1, u1, s1) == (n2, u2, s2) and np.array_equal(nb1, nb2)
    assert np.array_equal(ph1 != 0, ph2 != 0)
    nz = ph2 != 0
    assert np.max(np.abs(ph1[nz] - ph2[nz]) / ph2[nz]) < 1e-8   # second step: the order of the additions feeds back
    np.testing.assert_allclose(x1, x2, rtol=0, atol=1e-10)


def test_deterministic_clumping_on_device(gpu_tables, routing):
    """type_of_clumping 3 generated on the device from the resident density (clumping_module.F90:327-363): the grid
    is bit-identical to the restatement's, and a pass + per-cell pass with it matches the oracle"""
    if routing != "auto":
        pytest.skip("runs once")
    p = make_problem(24, nsrc=4, seed=19, state="random", use_LLS=True, clumping="grid")
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    p["type_of_clumping"] = 3
    coef = (0.0319, 1.2041, 2.7519)
    avg = float(np.mean(p["ndens"], dtype=np.float64))
    o = setup_oracle(p, tables=gpu_tables)
    o.deterministic_clumping(*coef, avg)
    e = setup_gpu(p, tables=gpu_tables)
    e.set_clumping_from_density(*coef, avg)
    assert np.array_equal(e.clumping_grid, o.clumping_grid)
    assert not np.array_equal(e.clumping_grid, p["clumping_grid"])
    ro = o.evolve3D(DT)
    rg = e.evolve3D(0.0, DT)
    assert rg.niter == ro.niter and rg.total_updates == ro.total_updates
    np.testing.assert_allclose(e.xh, o.xh, rtol=0, atol=X_ATOL)
    assert rg.final_stats.totrec == pytest.approx(ro.final_stats.totrec, rel=1e-6)
    e.close()
