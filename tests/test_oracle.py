"""CPU tests: pins of the oracle (the reference ships no tests; SURVEY 8c lists the pins below)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from problems import make_problem, setup_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_constants_bit_exact():
    """the values gfortran stores for the reference's parameters."""
    c = O.constants()
    expect = {
        "pi": "0x1.921fb60000000p+1", "sigma_HI_at_ion_freq": "0x1.d0dba20000000p-58",
        "eth0": "0x1.b322d00000000p+3", "ev2k": "0x1.6aa7bc0000000p+13",
        "temph0": "0x1.34362ab1a0600p+17", "colh0": "0x1.00a4e0fd09257p-34",
        "ev2fr": "0x1.b7e6980000000p+47", "ion_freq_HI": "0x1.75dc5329c5c00p+51",
        "bb_MaxFreq": "0x1.d38832dfab080p+56", "two_pi_over_c_square": "0x1.081ca7cc3579bp-67",
        "mu": "0x1.38d4fe0000000p+0", "Mpc": "0x1.46be335a58000p+81",
        "convergence_fraction": "0x1.a36e2e0000000p-14", "minimum_fractional_change": "0x1.0624de0000000p-10",
        "minimum_fraction_of_atoms": "0x1.5798ee0000000p-27", "max_coldensh": "0x1.158e460000000p+64",
        "tau_photo_limit": "0x1.ad7f2a0000000p-24", "sqrt3": "0x1.bb67ae0000000p+0",
        "sqrt2": "0x1.6a09e60000000p+0", "dlogtau": "0x1.89374bc6a7efap-7",
        "xh_initial": "0x1.a36e2e0000000p-13",
    }
    for name, hexv in expect.items():
        assert float.hex(getattr(c, name)) == hexv, name
    assert c.H0 == pytest.approx(2.2683083702422782e-18, rel=1e-15)
    assert c.rho_crit_0 == pytest.approx(9.203466430166128e-30, rel=1e-15)
    assert c.abu_he == 0.07400000095367432 and c.abu_c == 7.099999947968172e-07


def test_tables_known_answers():
    """SURVEY Appendix D + thick(tau=0) == S_star by construction (radiation_sed_parameters.F90:184-186)."""
    thick, thin, d = O.rad_ini()
    assert thick[0] == pytest.approx(1e48, rel=1e-13)
    known = {0: (1.000000000000e48, 4.567323859246e47), 1000: (9.999999955571e47, 4.567323831315e47),
             1500: (9.955706997619e47, 4.539496300295e47), 1668: (6.545380303526e47, 2.526098635800e47),
             1700: (3.983534748389e47, 1.238543418359e47), 1800: (7.137208700526e45, 4.186999762595e44),
             2000: (7.481540301262e31, 8.918105777421e28)}
    for i, (tk, tn) in known.items():
        assert thick[i] == pytest.approx(tk, rel=2e-12)
        assert thin[i] == pytest.approx(tn, rel=2e-12)
    assert np.all(np.diff(thick) <= 0) and np.all(np.diff(thin) <= 0)
    assert np.all(thick > 0)
    # Romberg weights integrate constants exactly: sum = number of intervals (to float32 accuracy of b(k))
    assert sum(d.romw7) == pytest.approx(128.0, rel=1e-6)
    assert d.S_scaling == pytest.approx(2.6247620, rel=1e-6)


def test_tables_golden():
    g = np.load(os.path.join(GOLD, "tables.npz"))
    thick, thin, d = O.rad_ini()
    np.testing.assert_array_equal(thick[g["idx"]], g["thick"])
    np.testing.assert_array_equal(thin[g["idx"]], g["thin"])
    assert thick.sum() == g["thick_sum"] and thin.sum() == g["thin_sum"]
    np.testing.assert_array_equal(np.array(d.romw7), g["romw7"])


def _cinterp_setup(N=16):
    o = O.Oracle(N)
    cd = o.coldensh_out
    rng = np.random.default_rng(0)
    cd[...] = rng.uniform(1e16, 1e19, size=cd.shape)
    return o, cd


def test_cinterp_exact_cases():
    """column_density.f90: on-axis cells take the single upstream neighbour with path 1; first-neighbour
    diagonals get the sqrt2/sqrt3 factors; ties go z, then y, then x."""
    o, cd = _cinterp_setup()
    src = (8, 8, 8)
    c = O.constants()
    # on the +z axis: cdensi = c4 (ip,jp,km)
    v, path = o.cinterp((8, 8, 11), src)
    assert path == 1.0 and v == pytest.approx(cd[11 - 2, 8 - 1, 8 - 1], rel=5e-16)  # (c*w)/w
    v, path = o.cinterp((8, 5, 8), src)   # -y axis
    assert path == 1.0 and v == pytest.approx(cd[8 - 1, 5, 8 - 1], rel=5e-16)
    v, path = o.cinterp((12, 8, 8), src)  # +x axis
    assert path == 1.0 and v == pytest.approx(cd[8 - 1, 8 - 1, 12 - 2], rel=5e-16)
    # first neighbours
    v, path = o.cinterp((9, 9, 9), src)
    assert path == pytest.approx(np.sqrt(3.0), rel=1e-15)
    assert v == pytest.approx(c.sqrt3 * cd[7, 7, 7], rel=5e-16)
    v, path = o.cinterp((9, 8, 9), src)
    assert path == pytest.approx(np.sqrt(2.0), rel=1e-15)
    assert v == pytest.approx(c.sqrt2 * cd[7, 7, 7], rel=5e-16)
    v, path = o.cinterp((8, 7, 9), src)
    assert v == pytest.approx(c.sqrt2 * cd[7, 7, 7], rel=5e-16)
    # body diagonal further out: only (im,jm,km) has weight
    v, path = o.cinterp((11, 11, 11), src)
    assert v == pytest.approx(cd[9, 9, 9], rel=5e-16) and path == pytest.approx(np.sqrt(3.0), rel=1e-15)
    # generic z-dominant cell: weights 1-|di|/|dk| etc.
    pos = (10, 9, 12)  # d = (2,1,4)
    v, path = o.cinterp(pos, src)
    dx, dy = 1 - 2 / 4, 1 - 1 / 4
    s = [(1 - dx) * (1 - dy), (1 - dy) * dx, (1 - dx) * dy, dx * dy]
    cs = [cd[10, 7, 8], cd[10, 7, 9], cd[10, 8, 8], cd[10, 8, 9]]
    w = [si / max(0.6, ci * c.sigma_HI_at_ion_freq) for si, ci in zip(s, cs)]
    assert v == pytest.approx(sum(ci * wi for ci, wi in zip(cs, w)) / sum(w), rel=1e-14)
    assert path == pytest.approx(np.sqrt(1 + (4 + 1) / 16.0), rel=1e-15)


def test_cinterp_periodic_wrap():
    o, cd = _cinterp_setup(12)
    src = (2, 11, 6)
    # cell at unwrapped (0, 13, 6): offsets (-2,+2,0); upstream cells wrap to i=12|1, j=12|1
    v, path = o.cinterp((0, 13, 6), src)
    assert v == pytest.approx(cd[5, 11, 0], rel=5e-16)  # (im,jm,k)=(1,12,6) -> C index [k-1, j-1, i-1]
    assert path == pytest.approx(np.sqrt(2.0), rel=1e-15)


@pytest.mark.parametrize("case", [
    dict(N=20, nsrc=2, seed=3, state="random"),
    dict(N=21, nsrc=3, seed=5, state="ionized", use_LLS=False),
    dict(N=24, nsrc=3, seed=5, state="ionized", use_LLS=True),
    dict(N=(16, 20, 12), nsrc=3, seed=5, state="ionized", use_LLS=False),
])
def test_walk_order_independence(case):
    """serial evolve2D order == axes/planes/octants order == scrambled Chebyshev-shell order, bit for bit
    (this is what licenses the GPU wavefront)."""
    p = make_problem(**case)
    res = []
    for order in (0, 1, 2):
        o = setup_oracle(p)
        o.set_walk_order(order)
        o.xh_av[...] = p["xh"]
        o.set_rates_to_zero()
        r = o.pass_all_sources()
        res.append((o.phih.copy(), o.coldensh_out.copy(), r))
    for k in (1, 2):
        np.testing.assert_array_equal(res[0][0], res[k][0])
        np.testing.assert_array_equal(res[0][1], res[k][1])
        assert res[0][2].sum_nbox_all == res[k][2].sum_nbox_all
        assert res[0][2].updates == res[k][2].updates
        assert res[0][2].photon_loss_all == pytest.approx(res[k][2].photon_loss_all, rel=1e-13)


@pytest.mark.parametrize("N,nbox,side", [(10, 1, 10), (12, 1, 11), (21, 2, 21), (22, 2, 21), (16, 2, 16), (31, 3, 31)])
def test_walk_bounds(N, nbox, side):
    """passes and final box, including the quirk that on even meshes with (N/2-1) a multiple
    of 5 the layer at -N/2 is never traced (N=12, 22, 512...)."""
    p = make_problem(N, nsrc=1, seed=2, state="ionized", use_LLS=False, flux=1e12)
    o = setup_oracle(p)
    o.set_loss_fraction(0.0)
    o.xh_av[...] = p["xh"]
    o.set_rates_to_zero()
    r = o.do_source(1)
    assert r.nbox == nbox
    assert r.updates == side ** 3
    assert int(np.count_nonzero(o.coldensh_out)) == side ** 3


def test_zero_flux_source_is_not_traced():
    p = make_problem(12, nsrc=2, seed=4, state="ionized")
    p["normflux"][0] = 0.0
    o = setup_oracle(p)
    o.xh_av[...] = p["xh"]
    o.set_rates_to_zero()
    r = o.do_source(1)
    assert r.nbox == 0 and r.updates == 0 and r.photon_loss_src == 0.0


def test_cube_symmetry():
    """one source at the centre of an odd uniform box: phih is invariant under axis permutations
    and reflections to rounding (SURVEY 8c pin 3)."""
    N = 15
    p = make_problem(N, nsrc=1, seed=1, state="ionized", use_LLS=False, dens="uniform", srcpos=[[8, 8, 8]])
    o = setup_oracle(p)
    o.xh_av[...] = p["xh"]
    o.set_rates_to_zero()
    o.pass_all_sources()
    ph = o.phih.copy()
    # not bit-exact: xc=alam*di+real(i0) rounds differently for +di and -di (column_density.f90:110)
    for mirrored in (ph[::-1, :, :], ph[:, ::-1, :], ph[:, :, ::-1]):
        np.testing.assert_allclose(ph, mirrored, rtol=1e-12, atol=0)
    for perm in ((1, 0, 2), (2, 1, 0), (0, 2, 1), (1, 2, 0)):
        np.testing.assert_allclose(ph, ph.transpose(perm), rtol=1e-13, atol=0)


def test_photoion_rates_photon_conservation():
    """Gamma_cell*vol = Gamma_in - Gamma_out in both branches; thin branch used below 1e-7 in tau."""
    o = O.Oracle(8)
    c = O.constants()
    vol = 3.0e70
    for n_in, dn in ((1e17, 1e16), (1e15, 1e17), (3e18, 2e18)):
        cell, pin, pout = o.photoion_rates(n_in, n_in + dn, vol, 2.0)
        assert cell * vol == pytest.approx(pin - pout, rel=1e-12)
        assert pin > pout > 0
    n_in, dn = 1e12, 1e9   # delta tau = 6.3e-9 < 1e-7
    cell, pin, pout = o.photoion_rates(n_in, n_in + dn, vol, 2.0)
    thick, thin, _ = O.rad_ini()
    assert cell * vol == pytest.approx(2.0 * (dn * c.sigma_HI_at_ion_freq) * thin[1000], rel=2e-2)
    assert cell * vol == pytest.approx(pin - pout, rel=1e-6)
    # below the table: tau < 1e-20 clamps to entry 1; zero flux gives zero
    assert o.photoion_rates(0.0, 1e16, vol, 1.0)[1] == pytest.approx(thick[0], rel=1e-12)
    assert o.photoion_rates(1e17, 2e17, vol, 0.0) == (0.0, 0.0, 0.0)


def test_doric_matches_ode():
    """doric is the closed-form solution of dx1/dt = aih0*(1-x1) - ne*brech0*x1 at fixed ne (SURVEY 8c pin 7)."""
    from scipy.integrate import solve_ivp
    o = O.Oracle(8)
    c = O.constants()
    T, ne, nh, phih, dt = 1e4, 1e-4, 2e-4, 3e-13, 3e13
    brech0 = 1.0 * c.bh00 * (T / 1e4) ** c.albpow
    acolh0 = c.colh0 * np.sqrt(T) * np.exp(-c.temph0 / T)
    aih0 = phih + ne * acolh0
    x_old = 0.2
    xf, xav = o.doric(dt, T, ne, nh, [1 - x_old, x_old], [1 - x_old, x_old], phih)
    sol = solve_ivp(lambda t, y: [aih0 * (1 - y[0]) - ne * brech0 * y[0], y[0]], (0, dt), [x_old, 0.0],
                    rtol=1e-12, atol=1e-14, method="DOP853")
    assert xf[1] == pytest.approx(sol.y[0, -1], rel=1e-9)
    assert xav[1] == pytest.approx(sol.y[1, -1] / dt, rel=1e-9)
    assert xf[0] + xf[1] == pytest.approx(1.0, abs=1e-15)


def test_stromgren_sphere():
    """Non-cosmological, no LLS, C=1, uniform gas, one source: ionized volume follows
    V(t) = V_S (1 - exp(-t/t_rec)) (SURVEY 8c pin 6), and the photon-conservation number stays near 1."""
    N = 31
    c = O.constants()
    n_H = 1e-3
    dr = 1.5e21
    ndot = 5e48
    p = make_problem(N, nsrc=1, seed=1, state="neutral", use_LLS=False, dens="uniform", srcpos=[[16, 16, 16]])
    p["ndens"][...] = n_H
    p["dr"] = np.array([dr] * 3)
    p["vol"] = dr ** 3
    p["normflux"] = np.array([ndot / 1e48])
    o = setup_oracle(p)
    alpha = c.bh00
    t_rec = 1.0 / (alpha * n_H)
    V_S = ndot / (alpha * n_H ** 2)
    dt = 0.05 * t_rec
    t = 0.0
    for step in range(10):
        rep = o.evolve3D(dt)
        t += dt
        assert rep.converged == 1
        assert abs(rep.final_stats.photcons - 1.0) < 0.05
    V_num = (o.xh.sum() - c.xh_initial * N ** 3) * dr ** 3
    V_ana = V_S * (1 - np.exp(-t / t_rec))
    assert V_num == pytest.approx(V_ana, rel=0.06)
    r_cells = (3 * V_ana / (4 * np.pi)) ** (1 / 3) / dr
    assert 3 < r_cells < N / 2 - 2  # the front is inside the box, so the test is meaningful


@pytest.mark.parametrize("name", ["g16_lls", "g20_clump", "g12x16x10"])
def test_oracle_reproduces_golden(name):
    import sys
    sys.path.insert(0, GOLD)
    import make_golden as mg
    g = np.load(os.path.join(GOLD, name + ".npz"))
    p = mg.make_case(name)
    o = setup_oracle(p)
    o.xh_av[...] = p["xh"]
    o.set_rates_to_zero()
    r = o.pass_all_sources()
    np.testing.assert_array_equal(o.phih, g["phih_pass"])
    assert r.sum_nbox_all == g["sum_nbox"] and r.updates == g["updates"]
    o2 = setup_oracle(p)
    rep = o2.evolve3D(mg.DT)
    assert rep.niter == g["niter"]
    np.testing.assert_array_equal(o2.xh, g["xh"])
    np.testing.assert_array_equal(np.array(rep.conv_flag[:rep.niter + 1]), g["conv_flag"])
    assert rep.final_stats.photcons == g["photcons"]


def test_source_parallel_threads_match_serial():
    """the CPU-baseline mode (sources dealt to threads with private rate grids, summed afterwards: the
    MPI picture of master_slave.F90:85 + evolve.F90:599) gives the serial result up to summation order."""
    p = make_problem(20, nsrc=7, seed=21, state="random")
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    outs = []
    for nt in (1, 3):
        o = setup_oracle(p)
        o.set_threads(nt)
        o.xh_av[...] = p["xh"]
        o.set_rates_to_zero()
        r = o.pass_all_sources()
        outs.append((o.phih.copy(), r))
    np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=1e-12, atol=0)
    assert outs[0][1].updates == outs[1][1].updates and outs[0][1].sum_nbox_all == outs[1][1].sum_nbox_all


def _box_updates(mesh, nbox, subboxsize=5, max_subbox=1000):
    """closed form the product uses to count updates from a source's final nbox (SURVEY A2b;
    c2b_api.cu: box_updates): cells of [-min(5n,L), +min(5n,R)] per axis"""
    if nbox <= 0:
        return 0
    u = 1
    for m in mesh:
        R = min(max_subbox, m // 2 - 1 + m % 2)
        L = min(max_subbox, m // 2)
        u *= min(subboxsize * nbox, R) + min(subboxsize * nbox, L) + 1
    return u


@pytest.mark.parametrize("mesh", [(10, 10, 10), (12, 12, 12), (21, 21, 21), (22, 22, 22), (16, 20, 12), (24, 14, 30), (31, 18, 18)])
def test_update_count_closed_form(mesh):
    """the number of evolve0D calls passing the gate equals the volume of the final subbox, for every subbox
    count a trace can end with (forced here through the loss threshold), cubic and non-cubic meshes"""
    p = make_problem(mesh, nsrc=1, seed=9, state="ionized", use_LLS=True, flux=1e10)
    seen = set()
    for lf in (0.0, 1e-6, 1e-3, 3e-2, 0.2, 0.6, 0.9, 0.999):
        o = setup_oracle(p)
        o.set_loss_fraction(lf)
        o.xh_av[...] = p["xh"]
        o.set_rates_to_zero()
        r = o.do_source(1)
        assert r.updates == _box_updates(mesh, r.nbox)
        assert int(np.count_nonzero(o.coldensh_out)) == r.updates
        seen.add(r.nbox)
    if min(mesh) > 12:
        assert len(seen) >= 2


def test_upstream_cells_outside_previous_shell_have_zero_weight():
    """SURVEY A3, the fact the GPU wavefront rests on: the upstream cells cinterp reads that are NOT in the
    previous Chebyshev shell (same-shell cells on cube edges/diagonals, the i-1 neighbour of an on-axis cell)
    carry a bilinear weight of exactly 0.  Checked by poisoning every cell outside shell r-1 with 1e300: the
    interpolated column density must not change by a single bit.  Radii up to 99, sources near a corner so
    that the periodic wrap is exercised."""
    N = 200
    rng = np.random.default_rng(5)
    zero = O.Oracle(N)
    pois = O.Oracle(N)
    pois.coldensh_out[...] = 1e300
    cz, cp = zero.coldensh_out, pois.coldensh_out
    checked = 0
    for src in ((3, 198, 100), (100, 100, 100), (200, 1, 2)):
        s = np.array(src)
        for r in list(range(1, 12)) + [17, 25, 40, 64, 77, 99]:
            # cells of shell r: corners, edges, face centres (on-axis), near-diagonals and random ones
            cells = set()
            for sg in ((1, 1, 1), (-1, 1, -1), (1, -1, -1), (-1, -1, 1)):
                cells.add((sg[0] * r, sg[1] * r, sg[2] * r))
                cells.add((sg[0] * r, sg[1] * r, 0))
                cells.add((sg[0] * r, 0, sg[2] * r))
                cells.add((0, sg[1] * r, sg[2] * r))
                cells.add((sg[0] * r, 0, 0))
                cells.add((0, sg[1] * r, 0))
                cells.add((0, 0, sg[2] * r))
                cells.add((sg[0] * r, sg[1] * (r - 1), sg[2] * 1))
                cells.add((sg[0] * 1, sg[1] * r, sg[2] * (r - 1)))
            for _ in range(12):
                d = rng.integers(-r, r + 1, size=3)
                d[rng.integers(0, 3)] = r * rng.choice((-1, 1))
                cells.add(tuple(int(v) for v in d))
            for d in cells:
                pos = s + np.array(d)
                # the (up to) four upstream cells, one step closer to the source along every axis
                sg = [1 if v >= 0 else -1 for v in d]
                ups = set()
                for ox in (0, 1):
                    for oy in (0, 1):
                        for oz in (0, 1):
                            ups.add((d[0] - ox * sg[0], d[1] - oy * sg[1], d[2] - oz * sg[2]))
                touched = []
                for u in ups:
                    if max(abs(u[0]), abs(u[1]), abs(u[2])) == r - 1:      # in the previous shell: a real value
                        q = (s + np.array(u) - 1) % N
                        v = float(rng.uniform(1e16, 3e19))
                        cz[q[2], q[1], q[0]] = v
                        cp[q[2], q[1], q[0]] = v
                        touched.append(q)
                a = zero.cinterp(pos, s)
                b = pois.cinterp(pos, s)
                assert a == b, (src, d)
                for q in touched:
                    cz[q[2], q[1], q[0]] = 0.0
                    cp[q[2], q[1], q[0]] = 1e300
                checked += 1
    assert checked > 1500


def test_omp_in_source_mode_is_exact():
    """the reference's OpenMP build (evolve_source.F90:141-186: all threads inside one source, 6 axes / 12 planes /
    8 octants with a barrier after each group) gives bit-identical rates: the sweeps of a group touch disjoint cells"""
    p = make_problem(20, nsrc=3, seed=31, state="random", use_LLS=True)
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    a = setup_oracle(p)
    a.xh_av[...] = p["xh"]
    a.set_rates_to_zero()
    ra = a.pass_all_sources()
    b = setup_oracle(p)
    b.set_threads(4)
    b.set_omp_in_source(True)
    b.xh_av[...] = p["xh"]
    b.set_rates_to_zero()
    rb = b.pass_all_sources()
    assert (ra.updates, ra.sum_nbox_all) == (rb.updates, rb.sum_nbox_all)
    assert np.array_equal(a.phih, b.phih)
    assert rb.photon_loss_all == pytest.approx(ra.photon_loss_all, rel=1e-12)


def test_deterministic_clumping_model():
    """deterministic_clumping (clumping_module.F90:327-363): the quadratic fit in ndens/avg_dens, formed in double in
    the Fortran order of operations and stored as default real"""
    from problems import make_problem
    p = make_problem(12, nsrc=1, seed=2)
    o = O.Oracle(12)
    o.set_density(p["ndens"])
    p1, p2, p3 = 0.0319, 1.2041, 2.7519
    avg = float(np.mean(p["ndens"], dtype=np.float64))
    o.deterministic_clumping(p1, p2, p3, avg)
    nd = p["ndens"].astype(np.float64)
    want = (p1 * nd / avg * nd / avg + p2 * nd / avg + p3).astype(np.float32)
    assert np.array_equal(o.clumping_grid, want)
    assert o.clumping_grid.dtype == np.float32 and float(o.clumping_grid.min()) > p3
