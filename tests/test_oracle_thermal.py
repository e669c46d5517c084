"""CPU pins of the non-isothermal restatement (thermal.f90, cooling.f90, heat_lookuptable, the heating tables):
the reference cannot run this path as shipped (tables/corocool.tab is absent, c2ray_parameters.f90:28), so these
are closed-form and independent-quadrature checks of the restatement itself."""
import numpy as np
import pytest

from problems import make_problem
from thermal_common import cooling_table, setup_thermal_oracle
from c2ray3dm_b200 import constants as K
from oracle import oracle as O

YEAR = 3.15576e7


@pytest.fixture(scope="module")
def tables4():
    return O.rad_ini_heat()


def test_heat_tables_against_independent_quadrature(tables4):
    """heat_thick(tau) = int hplanck*(nu-nu_HI)*SED*exp(-tau*sigma) dnu: compared with a fine trapezoid rule on the
    same SED; the ratio to the photon table at tau=0 is the mean excess energy of a 5e4 K black body"""
    thick, thin, hthick, hthin = tables4
    assert np.array_equal(O.rad_ini()[0], thick)          # same photon tables with and without the heat tables
    nu0, nu1 = K.bb_MinFreq, K.bb_MaxFreq
    nu = np.linspace(nu0, nu1, 400001)
    x = nu * K.hplanck / (K.k_B * K.bb_Teff)
    sed = np.where(x < 700, nu * nu / np.expm1(np.minimum(x, 700)), 0.0)
    cs = (nu / nu0) ** (-K.pl_index_cross_section_HI)
    for it in (0, 1500, 1668, 1800):
        tau = 0.0 if it == 0 else 10.0 ** (-20.0 + 0.012 * (it - 1))
        num = np.trapezoid(K.hplanck * (nu - nu0) * sed * np.exp(-tau * cs), nu)
        den = np.trapezoid(sed * np.exp(-tau * cs), nu)
        assert hthick[it] / thick[it] == pytest.approx(num / den, rel=2e-6)
        num2 = np.trapezoid(K.hplanck * (nu - nu0) * sed * cs * np.exp(-tau * cs), nu)
        den2 = np.trapezoid(sed * cs * np.exp(-tau * cs), nu)
        assert hthin[it] / thin[it] == pytest.approx(num2 / den2, rel=2e-6)
    ev = hthick[0] / thick[0] / 1.602e-12
    assert 6.0 < ev < 7.5                                  # mean excess energy per photo-ionization, eV
    assert np.all(np.diff(hthick) <= 0)                    # hardening spectrum: monotone in tau


def test_heating_rate_is_photon_rate_times_excess_energy(tables4):
    """heat per photo-ionization = mean excess energy of the absorbed photons: never below the optically thin
    value at tau=0 (heat_thin(0)/thin(0)), growing along a ray as the spectrum hardens, and bounded by the
    value of the hardened spectrum at the largest optical depth in the box"""
    p = make_problem(12, nsrc=1, seed=1, state="ionized", use_LLS=False, srcpos=[[6, 6, 6]])
    o = setup_thermal_oracle(p, tables4)
    o.xh_av[...] = p["xh"]
    o.set_rates_to_zero()
    o.do_source(1)
    ph, hh = o.phih, o.phiheat
    nz = ph > 0
    assert np.array_equal(hh > 0, nz)
    nHI = (1.0 - p["xh"]) * p["ndens"].astype(np.float64)
    ratio = hh / np.where(nz, ph * nHI, 1.0)               # erg per photo-ionization
    thin_excess = tables4[3] / tables4[1]                  # as a function of the table optical depth
    tau_max = float(o.coldensh_out.max()) * K.sigma_HI_at_ion_freq
    imax = int(1 + (np.log10(tau_max) + 20.0) / 0.012) + 1
    assert np.all(ratio[nz] >= thin_excess[0] * (1 - 1e-9))
    assert np.all(ratio[nz] <= thin_excess[imax] * (1 + 1e-9))
    ray = ratio[5, 5, 5:11]                                # along +x from the source cell (0-based 5,5,5)
    assert np.all(np.diff(ray) > 0)


def test_thermal_loses_energy_without_heating(tables4):
    """no sources: heating = 0, so the internal energy (n + n_e) k T / (gamma-1) of every cell drops (the
    temperature itself may rise while the gas recombines: fewer particles share the energy); cosmological
    cooling adds to the radiative one; at convergence current = intermed"""
    p = make_problem(8, nsrc=1, seed=2, state="ionized", use_LLS=False)
    p["normflux"][:] = 0.0
    nd = p["ndens"].astype(np.float64)
    e0 = (nd + nd * (p["xh"] + K.abu_c)) * 2.0e4
    res = []
    for cosmo in (False, True):
        o = setup_thermal_oracle(p, tables4, cosmological=cosmo, T0=2.0e4)
        r = o.evolve3D(1e6 * YEAR)
        T = o.temperature_grid.astype(np.float64)
        assert r.converged == 1
        e1 = (nd + nd * (o.xh + K.abu_c)) * T[..., 0]
        assert np.all(e1 < e0) and np.all(T[..., 0] >= 1.0)
        assert np.array_equal(T[..., 0], T[..., 2])            # set_final_temperature_point
        res.append(float(e1.sum()))
    assert res[1] < res[0]


def test_energy_equation_against_a_fine_ode_solve(tables4):
    """one uniform cell population, fixed ionization (neutral fraction ~ epsilon via a large rate is avoided: use no
    sources and a fully ionized start with recombination negligible over a short dt): T(t) from thermal() vs a
    4th-order Runge-Kutta solve of dE/dt = -Lambda(T) n n_e with the same table interpolation"""
    p = make_problem(6, nsrc=1, seed=3, state="ionized", use_LLS=False, dens="uniform")
    p["normflux"][:] = 0.0
    p["xh"][...] = 1.0 - 1e-5
    p["ndens"] = (p["ndens"] * 1000.0).astype(np.float32)   # a dense cell: cooling time ~ 10 Myr
    dt = 2e6 * YEAR
    o = setup_thermal_oracle(p, tables4, cosmological=False, T0=3.0e4)
    o.evolve3D(dt)
    T_end = float(o.temperature_grid[0, 0, 0, 0])
    # independent solve (x changes by < 1e-4 over dt at this density, so n_e is constant to that accuracy)
    lt, lc = cooling_table()
    n = float(p["ndens"][0, 0, 0])
    xav = float(o.xh_av[0, 0, 0])
    ne = n * (xav + K.abu_c)

    def lam(T):
        tpos = (np.log10(T) - lt[0]) / (lt[1] - lt[0])
        i = int(min(59, max(0, np.floor(tpos))))
        c0, c1 = 10.0 ** lc[i], 10.0 ** lc[i + 1]
        return c0 + (c1 - c0) * (tpos - i)

    kB = 1.381e-16
    E = (n + n * (1.0 - 1e-5 + K.abu_c)) * kB * 3.0e4 / (2.0 / 3.0)
    nsub = 20000
    h = dt / nsub

    def f(E_):
        T_ = E_ * (2.0 / 3.0) / (kB * (n + ne))
        return -n * ne * lam(T_)
    for _ in range(nsub):
        k1 = f(E); k2 = f(E + 0.5 * h * k1); k3 = f(E + 0.5 * h * k2); k4 = f(E + h * k3)
        E += h * (k1 + 2 * k2 + 2 * k3 + k4) / 6.0
    T_ref = E * (2.0 / 3.0) / (kB * (n + n * (float(o.xh[0, 0, 0]) + K.abu_c)))
    # thermal() is first order with <= 10 % energy steps: a few per cent is what the scheme delivers
    assert T_end == pytest.approx(T_ref, rel=5e-2)
    E0 = (n + n * (1.0 - 1e-5 + K.abu_c)) * kB * 3.0e4 / (2.0 / 3.0)
    assert E < 0.9 * E0                                     # the cell did lose an appreciable part of its energy


def test_threaded_pass_matches_serial_with_heating(tables4):
    p = make_problem(16, nsrc=6, seed=8, state="random", use_LLS=True)
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    a = setup_thermal_oracle(p, tables4)
    b = setup_thermal_oracle(p, tables4)
    b.set_threads(3)
    ra, rb = a.evolve3D(1e6 * YEAR), b.evolve3D(1e6 * YEAR)
    assert ra.niter == rb.niter
    np.testing.assert_allclose(a.phiheat, b.phiheat, rtol=1e-6, atol=0)   # summation order differs between the thread counts
    np.testing.assert_allclose(a.temperature_grid, b.temperature_grid, rtol=1e-6)
    np.testing.assert_allclose(a.xh, b.xh, rtol=0, atol=1e-9)
