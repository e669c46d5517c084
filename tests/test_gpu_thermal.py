"""GPU parity of the non-isothermal path (isothermal=.false.: heat_lookuptable, phiheat_grid, thermal.f90,
cooling.f90, cosmo_cool, the 3 x float temperature grid) against the oracle, with a synthetic 61-row cooling
table through the ABI (the reference does not ship tables/corocool.tab).

Tolerances: ionized fractions abs 1e-6, temperatures rel 1e-6 (they are stored as default real), heating and
photo-ionization rates rel 1e-6."""
import numpy as np
import pytest

from problems import make_problem
from thermal_common import cooling_table, setup_thermal_gpu, setup_thermal_oracle

pytestmark = pytest.mark.gpu
YEAR = 3.15576e7
DT = 1e6 * YEAR


@pytest.fixture(scope="module")
def tables4():
    """photo and heat tables built on the device (rad_ini with isothermal=.false.), checked against the oracle's"""
    from oracle import oracle as O
    from problems import setup_gpu
    e = setup_gpu(make_problem(8), isothermal=0)
    thick, thin = e.rad_ini()
    hthick, hthin = e.heat_tables()
    e.close()
    o = O.rad_ini_heat()
    for g, c in zip((thick, thin, hthick, hthin), o):
        np.testing.assert_allclose(g, c, rtol=1e-12, atol=0)
    return thick, thin, hthick, hthin


def _close_rates(gpu, cpu, rtol=1e-6):
    gpu, cpu = np.asarray(gpu).reshape(-1), np.asarray(cpu).reshape(-1)
    nz = cpu != 0
    assert np.array_equal(gpu != 0, nz)
    assert float(np.max(np.abs(gpu[nz] - cpu[nz]) / np.abs(cpu[nz]))) <= rtol


CASES = [
    dict(N=32, nsrc=6, seed=61, state="random", use_LLS=True),
    dict(N=32, nsrc=5, seed=62, state="neutral", use_LLS=True, flux=3e8),
    dict(N=(24, 32, 20), nsrc=4, seed=63, state="random", use_LLS=True, type_of_LLS=2, clumping="grid"),
    dict(N=64, nsrc=20, seed=64, state="random", use_LLS=True, clumping="scalar2"),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "N%s_%s" % (c["N"], c["state"]))
@pytest.mark.parametrize("routing", ["cta", "cluster"])
def test_heating_pass_matches_oracle(case, routing, tables4, monkeypatch):
    """one pass_all_sources: phih_grid and phiheat_grid with each of the two ray-trace kernels"""
    if routing == "cta":
        monkeypatch.setenv("C2B_CLUSTER_MIN_NBOX", "100000")
    else:
        monkeypatch.setenv("C2B_CLUSTER_MIN_NBOX", "0")
        monkeypatch.setenv("C2B_CLUSTER_MAX_SOURCES", "1000000")
    p = make_problem(**case)
    if case["state"] == "random":
        p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    o = setup_thermal_oracle(p, tables4)
    o.xh_av[...] = p["xh"]
    o.set_rates_to_zero()
    ro = o.pass_all_sources()
    e = setup_thermal_gpu(p, tables4)
    e.begin_step()
    rg = e.pass_all_sources()
    assert rg.updates == ro.updates and rg.sum_nbox_all == ro.sum_nbox_all
    _close_rates(e.phih_grid, o.phih)
    _close_rates(e.phiheat_grid, o.phiheat)
    e.close()


@pytest.mark.parametrize("case", CASES, ids=lambda c: "N%s_%s" % (c["N"], c["state"]))
@pytest.mark.parametrize("routing", ["auto", "warp"])
def test_thermal_evolve3d_matches_oracle(case, routing, tables4, monkeypatch):
    """two consecutive evolve3D steps with heating and cooling: iteration counts and convergence counters exact,
    ionized fractions, the three temperatures, the rates and the photon statistics within tolerance ("warp": every
    source whose previous trace ended after one subbox goes to the one-warp-per-source kernel, heating variant)"""
    if routing == "warp":
        monkeypatch.setenv("C2B_CLUSTER_MIN_NBOX", "100000")
        monkeypatch.setenv("C2B_WARP_MIN_SOURCES", "1")
    else:
        monkeypatch.delenv("C2B_WARP_MIN_SOURCES", raising=False)
    p = make_problem(**case)
    if case["state"] == "random":
        p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    o = setup_thermal_oracle(p, tables4)
    e = setup_thermal_gpu(p, tables4)
    for step in range(2):
        ro = o.evolve3D(DT)
        rg = e.evolve3D(step * DT, DT)
        assert (rg.niter, rg.converged) == (ro.niter, ro.converged)
        assert list(rg.conv_flag[1:rg.niter + 1]) == list(ro.conv_flag[1:ro.niter + 1])
        assert rg.total_updates == ro.total_updates
        np.testing.assert_allclose(e.xh, o.xh, rtol=0, atol=1e-6)
        np.testing.assert_allclose(e.xh_av, o.xh_av, rtol=0, atol=1e-6)
        Tg, To = e.temperature_grid, o.temperature_grid
        np.testing.assert_allclose(Tg, To, rtol=1e-6, atol=0)
        assert np.array_equal(Tg[..., 0], Tg[..., 2])            # set_final_temperature_point at convergence
        _close_rates(e.phih_grid, o.phih)
        _close_rates(e.phiheat_grid, o.phiheat)
        for n in ("photcons", "totrec", "totcollisions", "total_photon_loss"):
            assert getattr(rg.final_stats, n) == pytest.approx(getattr(ro.final_stats, n), rel=1e-6), n
    assert float(To[..., 0].max()) > 1.02e4                       # the gas near the sources was heated
    e.close()


def test_temperature_grid_roundtrip_and_restart_records(tables4):
    """temperature_grid goes through the ABI as the reference stores it (3 default reals per cell); the two extra
    records of a non-isothermal iteration dump (phiheat_grid, temperature_grid, evolve.F90:314-317) round-trip"""
    p = make_problem(16, nsrc=3, seed=70, state="random", use_LLS=True)
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    e = setup_thermal_gpu(p, tables4)
    rng = np.random.default_rng(5)
    tg = rng.uniform(5e3, 3e4, size=p["shape"] + (3,)).astype(np.float32)
    e.set_temperature_grid(tg)
    np.testing.assert_array_equal(e.temperature_grid, tg)
    o = setup_thermal_oracle(p, tables4)
    o.temperature_grid[...] = tg
    ro = o.evolve3D(DT)
    rg = e.evolve3D(0.0, DT)
    assert rg.niter == ro.niter
    np.testing.assert_allclose(e.temperature_grid, o.temperature_grid, rtol=1e-6, atol=0)
    import ctypes as C
    n = e.ncell
    ph = np.empty(n)
    t2 = np.empty(3 * n, dtype=np.float32)
    dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
    assert e.L.c2b_get_iter_state_thermal(e.h, ph.ctypes.data_as(dp), t2.ctypes.data_as(fp)) == 0
    e2 = setup_thermal_gpu(p, tables4)
    assert e2.L.c2b_set_iter_state_thermal(e2.h, ph.ctypes.data_as(dp), t2.ctypes.data_as(fp)) == 0
    np.testing.assert_array_equal(e2.phiheat_grid.reshape(-1), ph)
    np.testing.assert_array_equal(e2.temperature_grid.reshape(-1), t2)
    e.close()
    e2.close()


def test_isothermal_handle_rejects_thermal_calls():
    from problems import setup_gpu
    from c2ray3dm_b200 import C2RayError
    e = setup_gpu(make_problem(8))
    with pytest.raises(C2RayError):
        e.set_cooling_table(*cooling_table())
    with pytest.raises(C2RayError):
        _ = e.temperature_grid
    e.close()
    e2 = setup_gpu(make_problem(8), isothermal=0)
    e2.rad_ini()
    with pytest.raises(C2RayError) as ei:
        e2.evolve3D(0.0, DT)
    assert "cooling table" in str(ei.value)
    e2.close()
