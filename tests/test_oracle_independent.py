"""A second, independent transcription of three formulas of the reference -- written in plain Python straight from the
Fortran text, not from oracle/c2ray_oracle.c -- compared with the C restatement on random inputs.  The reference
cannot be run here (no Fortran compiler), so this guards the restatement against transcription slips: two independent
readings of column_density.f90, doric.f90 and radiation_photoionrates.F90 must agree."""
import math

import numpy as np
import pytest

from oracle import oracle as O


def _sign(a, b):
    """Fortran sign(a,b) for integers: |a| with the sign of b, b == 0 counts as positive"""
    return abs(a) if b >= 0 else -abs(a)


def cinterp_py(coldensh_out, mesh, pos, srcpos, sigma, sqrt2, sqrt3):
    """column_density.f90:29-271 (cinterp) and :276-293 (weightf); coldensh_out[k-1][j-1][i-1], positions 1-based"""
    i, j, k = pos
    i0, j0, k0 = srcpos
    idel, jdel, kdel = i - i0, j - j0, k - k0                      # :78-80
    idela, jdela, kdela = abs(idel), abs(jdel), abs(kdel)          # :81-83
    sgni, sgnj, sgnk = _sign(1, idel), _sign(1, jdel), _sign(1, kdel)   # :86-91
    im, jm, km = i - sgni, j - sgnj, k - sgnk                      # :92-94
    di, dj, dk = float(idel), float(jdel), float(kdel)             # :95-97

    def cd(ii, jj, kk):                                            # modulo(x-1,mesh)+1, :125-129
        return coldensh_out[(kk - 1) % mesh[2]][(jj - 1) % mesh[1]][(ii - 1) % mesh[0]]

    def weightf(c):                                                # :290
        return 1.0 / max(0.6, c * sigma)

    if kdela >= jdela and kdela >= idela:                          # :108
        alam = (float(km - k0) + sgnk * 0.5) / dk                  # :112
        xc = alam * di + float(i0)
        yc = alam * dj + float(j0)
        dx = 2.0 * abs(xc - (float(im) + 0.5 * sgni))              # :117-118
        dy = 2.0 * abs(yc - (float(jm) + 0.5 * sgnj))
        s1, s2, s3, s4 = (1. - dx) * (1. - dy), (1. - dy) * dx, (1. - dx) * dy, dx * dy   # :120-123
        c1, c2, c3, c4 = cd(im, jm, km), cd(i, jm, km), cd(im, j, km), cd(i, j, km)       # :130-133
        diag = kdela == 1 and (idela == 1 or jdela == 1)           # :152
        both = idela == 1 and jdela == 1
        path = math.sqrt((di * di + dj * dj) / (dk * dk) + 1.0)    # :168
    elif jdela >= idela and jdela >= kdela:                        # :173
        alam = (float(jm - j0) + sgnj * 0.5) / dj
        zc = alam * dk + float(k0)
        xc = alam * di + float(i0)
        dz = 2.0 * abs(zc - (float(km) + 0.5 * sgnk))
        dx = 2.0 * abs(xc - (float(im) + 0.5 * sgni))
        s1, s2, s3, s4 = (1. - dx) * (1. - dz), (1. - dz) * dx, (1. - dx) * dz, dx * dz
        c1, c2, c3, c4 = cd(im, jm, km), cd(i, jm, km), cd(im, jm, k), cd(i, jm, k)
        diag = jdela == 1 and (idela == 1 or kdela == 1)
        both = idela == 1 and kdela == 1
        path = math.sqrt((di * di + dk * dk) / (dj * dj) + 1.0)
    else:                                                          # :226 (idela largest)
        alam = (float(im - i0) + sgni * 0.5) / di
        zc = alam * dk + float(k0)
        yc = alam * dj + float(j0)
        dz = 2.0 * abs(zc - (float(km) + 0.5 * sgnk))
        dy = 2.0 * abs(yc - (float(jm) + 0.5 * sgnj))
        s1, s2, s3, s4 = (1. - dz) * (1. - dy), (1. - dz) * dy, (1. - dy) * dz, dy * dz
        c1, c2, c3, c4 = cd(im, jm, km), cd(im, j, km), cd(im, jm, k), cd(im, j, k)
        diag = idela == 1 and (jdela == 1 or kdela == 1)
        both = jdela == 1 and kdela == 1
        path = math.sqrt(1.0 + (dj * dj + dk * dk) / (di * di))
    w1, w2, w3, w4 = s1 * weightf(c1), s2 * weightf(c2), s3 * weightf(c3), s4 * weightf(c4)
    cdensi = (c1 * w1 + c2 * w2 + c3 * w3 + c4 * w4) / (w1 + w2 + w3 + w4)
    if diag:
        cdensi = (sqrt3 if both else sqrt2) * cdensi
    return cdensi, path


@pytest.mark.parametrize("mesh", [(16, 16, 16), (12, 17, 10)], ids=["cubic", "non_cubic"])
def test_cinterp_second_transcription(mesh):
    o = O.Oracle(mesh)
    cd = o.coldensh_out
    rng = np.random.default_rng(5)
    cd[...] = 10.0 ** rng.uniform(15.0, 20.5, size=cd.shape)       # weightf on both sides of its 0.6 floor
    c = O.constants()
    n = 0
    for _ in range(4000):
        src = tuple(int(rng.integers(1, mesh[d] + 1)) for d in range(3))
        # destination within the half box, possibly outside [1,mesh] (periodic wrap)
        off = tuple(int(rng.integers(-(mesh[d] // 2), mesh[d] // 2)) for d in range(3))
        if off == (0, 0, 0):
            continue
        pos = tuple(src[d] + off[d] for d in range(3))
        v_c, p_c = o.cinterp(pos, src)
        v_p, p_p = cinterp_py(cd, mesh, pos, src, c.sigma_HI_at_ion_freq, c.sqrt2, c.sqrt3)
        assert p_c == pytest.approx(p_p, rel=4e-16), (pos, src)
        assert v_c == pytest.approx(v_p, rel=1e-14), (pos, src)
        n += 1
    assert n > 3900


def doric_py(dt, temp0, rhe, xfh, xfh_av, phih, clumping, c):
    """doric.f90:33-134 for hydrogen"""
    brech0 = clumping * c.bh00 * (temp0 / 1e4) ** c.albpow          # :74
    sqrtt0 = math.sqrt(temp0)                                      # :76
    acolh0 = c.colh0 * sqrtt0 * math.exp(-c.temph0 / temp0)        # :77
    aih0 = phih + rhe * acolh0                                     # :80
    delth = aih0 + rhe * brech0                                    # :83
    eqxfh1 = aih0 / delth                                          # :84
    eqxfh0 = rhe * brech0 / delth                                  # :85
    deltht = delth * dt                                            # :88
    ee = math.exp(-deltht)                                         # :89
    x1 = (xfh[1] - eqxfh1) * ee + eqxfh1                           # :90
    x0 = (xfh[0] - eqxfh0) * ee + eqxfh0                           # :91
    if x0 < c.epsilon:                                             # :97-100
        x0 = c.epsilon
        x1 = 1.0 - c.epsilon
    avg_factor = 1.0 if deltht < 1.0e-8 else (1.0 - ee) / deltht   # :104-108
    xav1 = eqxfh1 + (xfh[1] - eqxfh1) * avg_factor                 # :112
    xav0 = 1.0 - xav1                                              # :113
    if xav0 < c.epsilon:                                           # :119
        xav0 = c.epsilon
    return (x0, x1), (xav0, xav1)


def test_doric_second_transcription():
    o = O.Oracle(8)
    c = O.constants()
    rng = np.random.default_rng(11)
    for _ in range(2000):
        dt = 10.0 ** rng.uniform(10.0, 15.0)
        temp0 = 10.0 ** rng.uniform(2.0, 5.0)
        rhe = 10.0 ** rng.uniform(-8.0, -2.0)
        x1 = float(rng.uniform(1e-6, 1 - 1e-6))
        phih = 0.0 if rng.uniform() < 0.2 else 10.0 ** rng.uniform(-20.0, -10.0)
        clump = float(np.float32(rng.uniform(1.0, 30.0)))   # clumping is a default real (clumping_module.F90:17)
        xfh = np.array([1.0 - x1, x1])
        got, got_av = o.doric(dt, temp0, rhe, rhe, xfh, xfh.copy(), phih, clump)
        want, want_av = doric_py(dt, temp0, rhe, xfh, xfh, phih, clump, c)
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(got_av, want_av, rtol=1e-12, atol=1e-300)


def photoion_rates_py(colum_in, colum_out, vol, nflux, thick, thin, c, numtau):
    """radiation_photoionrates.F90:71-179 (photoion_rates), :184-208 (set_tau_table_positions), :212-228 (read_table),
    :233-317 (photo_lookuptable) for one stellar source and the single frequency band the reference uses"""
    def positions(tau):
        ltau = math.log10(max(1.0e-20, tau))                                        # :197
        odpos = min(float(numtau), max(0.0, 1.0 + (ltau - c.minlogtau) / c.dlogtau))   # :198-199
        ipos = int(odpos)                                                           # :200
        return ipos, odpos - float(ipos), min(numtau, ipos + 1)                     # :201-204

    def read_table(table, p):                                                       # :224-226
        ipos, res, ipos_p1 = p
        return table[ipos] + (table[ipos_p1] - table[ipos]) * res

    tau_in = colum_in * c.sigma_HI_at_ion_freq                                      # :109-111
    tau_out = colum_out * c.sigma_HI_at_ion_freq                                    # :114-116
    p_in, p_out = positions(tau_in), positions(tau_out)
    photo_in = photo_out = photo_cell = 0.0
    if nflux > 0.0:                                                                 # :126
        phi_in = nflux * read_table(thick, p_in)                                    # :281-283
        photo_in += phi_in
        if abs(tau_out - tau_in) > float(np.float32(1.0e-7)):   # tau_photo_limit = 1.0e-7: a default-real literal (:244)
            phi_out = nflux * read_table(thick, p_out)                              # :293-295
            phi_all = phi_in - phi_out
        else:
            phi_all = nflux * (tau_out - tau_in) * read_table(thin, p_in)           # :301-304
            phi_out = phi_in - phi_all
        photo_out += phi_out
        photo_cell += phi_all / vol                                                 # :314-315
    return photo_cell, photo_in, photo_out


def test_photoion_rates_second_transcription():
    thick, thin, _ = O.rad_ini()
    o = O.Oracle(8)
    o.set_tables(thick, thin)
    c = O.constants()
    rng = np.random.default_rng(21)
    assert c.tau_photo_limit == float(np.float32(1.0e-7))     # the literal 1.0e-7 is a default real
    for n in range(6000):
        colum_in = 0.0 if n % 50 == 0 else 10.0 ** rng.uniform(10.0, 23.0)
        kind = n % 3
        if kind == 0:      # optically thick cell
            colum_out = colum_in + 10.0 ** rng.uniform(12.0, 22.0)
        elif kind == 1:    # optically thin cell (below tau_photo_limit)
            colum_out = colum_in + 10.0 ** rng.uniform(5.0, 10.0)
        else:              # around the switch
            colum_out = colum_in + 10.0 ** rng.uniform(10.0, 11.5)
        vol = 10.0 ** rng.uniform(60.0, 75.0)
        nflux = 0.0 if n % 97 == 0 else 10.0 ** rng.uniform(-3.0, 9.0)
        got = o.photoion_rates(colum_in, colum_out, vol, nflux)
        want = photoion_rates_py(colum_in, colum_out, vol, nflux, thick, thin, c, len(thick) - 1)
        for g, w in zip(got, want):
            assert g == pytest.approx(w, rel=1e-13, abs=0.0), (n, colum_in, colum_out)


def heat_lookup_py(colum_in, colum_out, vol, nflux, hthick, hthin, c, numtau):
    """heat_lookuptable (radiation_photoionrates.F90:323-417) as photoion_rates calls it for isothermal=.false. (:136-147)"""
    def positions(tau):                                                             # set_tau_table_positions, :184-208
        odpos = min(float(numtau), max(0.0, 1.0 + (math.log10(max(1.0e-20, tau)) - c.minlogtau) / c.dlogtau))
        ipos = int(odpos)
        return ipos, odpos - float(ipos), min(numtau, ipos + 1)

    def read_table(table, q):
        return table[q[0]] + (table[q[2]] - table[q[0]]) * q[1]

    tau_in, tau_out = colum_in * c.sigma_HI_at_ion_freq, colum_out * c.sigma_HI_at_ion_freq
    tau_cell = (colum_out - colum_in) * c.sigma_HI_at_ion_freq                      # :138-140 (colum_cell_HI, :106)
    if not nflux > 0.0:                                                             # :143
        return 0.0
    heat_in = nflux * read_table(hthick, positions(tau_in))                         # :377-378
    if abs(tau_out - tau_in) > float(np.float32(1.0e-4)):                           # tau_heat_limit, :333 (default real)
        heat_out = nflux * read_table(hthick, positions(tau_out))                   # :383-384
        return (heat_in - heat_out) / vol                                           # :385
    return nflux * tau_cell * read_table(hthin, positions(tau_in)) / vol            # :390-393


def do_source_py(p, ns, thick, thin, c, heat=None):
    """do_source (evolve_source.F90:58-221), serial branch with evolve2D (:227-267), and evolve0D
    (evolve_point.F90:83-299) for the isothermal path with a homogeneous LLS column (type_of_LLS 1) or none;
    periodic boundaries.  Returns (coldensh_out, phih_grid, nbox, photon_loss_src, number of evolve0D updates)."""
    mesh = p["mesh"]
    ndens, xh_av = p["ndens"], p["xh"]
    dr, vol = p["dr"], p["vol"]
    src = [int(v) for v in p["srcpos"][ns - 1]]
    nflux = float(p["normflux"][ns - 1])
    subboxsize, max_subbox = 5, 1000                                  # c2ray_parameters.f90:54,61
    max_coldensh = float(np.float32(2e19))                            # evolve_point.F90:95 (a default-real literal)
    coldensh_out = np.zeros((mesh[2], mesh[1], mesh[0]))              # evolve_source.F90:90
    phih = np.zeros_like(coldensh_out)
    phiheat = np.zeros_like(coldensh_out)
    lastpos_r = [src[d] + min(max_subbox, mesh[d] // 2 - 1 + mesh[d] % 2) for d in range(3)]   # :100
    lastpos_l = [src[d] - min(max_subbox, mesh[d] // 2) for d in range(3)]                       # :101
    state = {"loss": 0.0, "updates": 0}

    def evolve0d(rtpos, last_l, last_r):
        pos = [(rtpos[d] - 1) % mesh[d] for d in range(3)]            # 0-based modulo(rtpos-1,mesh), evolve_point.F90:122-124
        idx = (pos[2], pos[1], pos[0])
        if coldensh_out[idx] != 0.0:                                  # :128
            return
        state["updates"] += 1
        h_av1 = max(float(xh_av[idx]), c.epsilon)                     # :137
        h_av0 = max(1.0 - h_av1, c.epsilon)                           # :140
        ndens_p = float(ndens[idx])                                   # :146
        stop = False
        if rtpos == src:                                              # :151-160
            coldensh_in = 0.0
            path = 0.5 * dr[0]
            vol_ph = dr[0] * dr[1] * dr[2]
        else:
            coldensh_in, path = cinterp_py(coldensh_out, mesh, tuple(rtpos), tuple(src), c.sigma_HI_at_ion_freq,
                                           c.sqrt2, c.sqrt3)          # :165
            path = path * dr[0]                                       # :167
            xs = dr[0] * float(rtpos[0] - src[0])                     # :170-172
            ys = dr[1] * float(rtpos[1] - src[1])
            zs = dr[2] * float(rtpos[2] - src[2])
            dist2 = xs * xs + ys * ys + zs * zs                       # :173
            vol_ph = 4.0 * c.pi * dist2 * path                        # :177
            if p["use_LLS"]:                                          # :186-197, type_of_LLS == 1
                coldensh_in = coldensh_in + p["coldensh_LLS"] * path / dr[0]
        if coldensh_in > max_coldensh:                                # :201
            stop = True
        coldensh_out[idx] = coldensh_in + h_av0 * ndens_p * path      # :247-248 with coldens, doric.f90:153
        photo_out = 0.0
        if not stop:                                                  # :254-262
            cell, _, photo_out = photoion_rates_py(coldensh_in, coldensh_out[idx], vol_ph, nflux, thick, thin, c,
                                                   len(thick) - 1)
            phih[idx] += cell / (h_av0 * ndens_p)                     # :262, :283-284
            if heat is not None:                                      # :285-286: not divided by the neutral density
                phiheat[idx] += heat_lookup_py(coldensh_in, coldensh_out[idx], vol_ph, nflux, heat[0], heat[1], c,
                                               len(thick) - 1)
        if any(rtpos[d] == last_l[d] for d in range(3)) or any(rtpos[d] == last_r[d] for d in range(3)):   # :290-291
            state["loss"] += photo_out * vol / vol_ph                 # :292-293

    nbox = 0
    total_source_flux = nflux * p["S_star"]                           # evolve_source.F90:119
    photon_loss_src = total_source_flux                               # :121
    last_r, last_l = list(src), list(src)                             # :122-123
    while photon_loss_src > c.loss_fraction * total_source_flux and last_r[2] < lastpos_r[2] and \
            last_l[2] > lastpos_l[2]:                                 # :128-131
        nbox += 1
        state["loss"] = 0.0
        last_r = [min(src[d] + subboxsize * nbox, lastpos_r[d]) for d in range(3)]   # :135
        last_l = [max(src[d] - subboxsize * nbox, lastpos_l[d]) for d in range(3)]   # :136
        ks = list(range(src[2], last_r[2] + 1)) + list(range(src[2] - 1, last_l[2] - 1, -1))   # :190-199
        for k in ks:
            js = list(range(src[1], last_r[1] + 1)) + list(range(src[1] - 1, last_l[1] - 1, -1))   # evolve2D, :241-263
            for j in js:
                for i in list(range(src[0], last_r[0] + 1)) + list(range(src[0] - 1, last_l[0] - 1, -1)):
                    evolve0d([i, j, k], last_l, last_r)
        photon_loss_src = state["loss"]                               # :206
    if heat is not None:
        return coldensh_out, phih, nbox, photon_loss_src, state["updates"], phiheat
    return coldensh_out, phih, nbox, photon_loss_src, state["updates"]


@pytest.mark.parametrize("case", [dict(N=14, seed=3, state="random", use_LLS=True),
                                  dict(N=13, seed=4, state="random", use_LLS=False),
                                  dict(N=(12, 15, 10), seed=6, state="ionized", use_LLS=True)],
                         ids=["even_lls", "odd", "non_cubic_ionized"])
def test_do_source_second_transcription(case):
    """a whole single-source trace: the serial sweep, the growing subbox with its loss cut-off, the periodic half box
    (incl. the extra layer on the negative side of an even mesh), evolve0D with cinterp and the rate look-up"""
    from problems import make_problem, setup_oracle
    p = make_problem(case["N"], nsrc=3, seed=case["seed"], state=case["state"], use_LLS=case["use_LLS"], flux=3e7)
    if case["state"] == "random":
        p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    thick, thin, _ = O.rad_ini()
    c = O.constants()
    o = setup_oracle(p)
    o.xh_av[...] = p["xh"]
    for ns in (1, 2, 3):
        o.set_rates_to_zero()
        o.coldensh_out[...] = 0.0
        r = o.do_source(ns)
        cd, ph, nbox, loss, upd = do_source_py(p, ns, thick, thin, c)
        assert (r.nbox, r.updates) == (nbox, upd)
        assert np.array_equal(o.coldensh_out != 0, cd != 0)
        np.testing.assert_allclose(o.coldensh_out, cd, rtol=1e-13, atol=0)
        assert np.array_equal(o.phih != 0, ph != 0)
        np.testing.assert_allclose(o.phih, ph, rtol=1e-10, atol=0)   # Gamma_in - Gamma_out cancels at small dtau
        assert r.photon_loss_src == pytest.approx(loss, rel=1e-10)


def global_pass_py(p, xh, xh_av, xh_intermed, phih, dt, c, temper):
    """global_pass's loop (evolve.F90:548-555) over evolve0D_global (evolve_point.F90:305-406) and do_chemistry
    (:410-555), isothermal; electrondens is tped.f90:75-83.  Returns (conv_flag, new xh_intermed, new xh_av)."""
    new_int, new_av = xh_intermed.copy(), xh_av.copy()
    conv_flag = 0
    eps, mfc, mfa = c.epsilon, c.minimum_fractional_change, c.minimum_fraction_of_atoms
    for idx in np.ndindex(xh.shape):
        h_old1 = max(eps, float(xh[idx]))                          # evolve_point.F90:349
        h_av1 = max(eps, float(xh_av[idx]))                        # :350
        h_old0 = 1.0 - h_old1                                      # :352
        h_av0 = 1.0 - h_av1                                        # :353
        ndens_p = float(p["ndens"][idx])                           # :357
        ph = float(phih[idx])                                      # :363
        clump = float(p["clumping_grid"][idx]) if p["clumping_grid"] is not None else float(np.float32(p["clumping"]))   # :441-443
        nit = 0
        while True:                                                # :448
            nit += 1
            yh0_av_old = h_av0                                     # :453
            de = ndens_p * (h_av1 + c.abu_c)                       # :461, tped.f90:81
            (h0, h1), (h_av0, h_av1) = doric_py(dt, temper, de, (h_old0, h_old1), None, ph, clump, c)   # :457, :513
            if abs((h_av0 - yh0_av_old) / h_av0) < mfc or h_av0 < mfa:   # :530-535 (the temperature term is 0 < mfc)
                break
            if nit > 400:                                          # :540
                break
        yh1_prev = max(eps, float(xh_av[idx]))                     # :376
        yh0_prev = 1.0 - yh1_prev                                  # :377
        if abs(h_av0 - yh0_prev) > mfc and abs((h_av0 - yh0_prev) / h_av0) > mfc and h_av0 > mfa:   # :382-384
            conv_flag += 1                                         # :389
        new_int[idx] = h1                                          # :400
        new_av[idx] = h_av1                                        # :401
    return conv_flag, new_int, new_av


@pytest.mark.parametrize("clumping", ["scalar", "scalar2", "grid"])
def test_global_pass_second_transcription(clumping):
    from problems import make_problem, setup_oracle
    p = make_problem(10, nsrc=3, seed=8, state="random", use_LLS=True, clumping=clumping, flux=3e7)
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    c = O.constants()
    o = setup_oracle(p)
    o.xh_av[...] = p["xh"]
    o.xh_intermed[...] = p["xh"]
    o.state_before()
    o.set_rates_to_zero()
    r = o.pass_all_sources()
    phih = o.phih.copy()
    dt = 1e6 * c.YEAR
    for it in range(2):    # two per-cell passes with the same rates: the second starts from the first one's xh_av
        xh_av_in, xh_int_in = o.xh_av.copy(), o.xh_intermed.copy()
        conv, want_int, want_av = global_pass_py(p, o.xh.copy(), xh_av_in, xh_int_in, phih, dt, c, p["temper"])
        g = o.global_pass(dt, r.photon_loss_all)
        assert g.conv_flag == conv
        np.testing.assert_allclose(o.xh_intermed, want_int, rtol=0, atol=1e-14)
        np.testing.assert_allclose(o.xh_av, want_av, rtol=0, atol=1e-14)
    assert conv < 1000


def photon_statistics_py(p, xh_before, xh, xh_av, photon_loss_all, dt, c):
    """state_before / state_after / total_rates / total_ionizations / report_photonstatistics
    (photonstatistics.F90:104-281) for the isothermal case; LLS_loss stays 0 in the reference (evolve0D passes the
    never-assigned phi%photo_in_HI to total_LLS_loss, evolve_point.F90:273-274)"""
    nd = p["ndens"].astype(np.float64)
    mesh = p["mesh"]
    vol = p["vol"]
    clump = p["clumping_grid"].astype(np.float64) if p["clumping_grid"] is not None else float(np.float32(p["clumping"]))
    T = p["temper"]
    st = {}
    st["h0_before"] = float(np.sum(nd * (1.0 - xh_before))) * vol            # :118-131
    st["h0_after"] = float(np.sum(nd * (1.0 - xh))) * vol                    # :204-216
    st["h1_after"] = float(np.sum(nd * xh)) * vol
    yh1, yh0 = xh_av, 1.0 - xh_av                                            # total_rates(dt,xh_av), :155-163
    ne = nd * (yh1 + c.abu_c)                                                # tped.f90:81
    st["totrec"] = float(np.sum(nd * yh1 * ne * clump * c.bh00 * (T / 1e4) ** c.albpow)) * vol * dt      # :170-172,:183
    st["totcollisions"] = float(np.sum(nd * yh0 * ne * c.colh0 * math.sqrt(T) * math.exp(-c.temph0 / T))) * vol * dt
    st["dh0"] = st["h0_before"] - st["h0_after"]                             # :224
    st["total_ion"] = st["totrec"] + st["dh0"]                               # :225
    m3 = float(np.float32(mesh[0]) * np.float32(mesh[1]) * np.float32(mesh[2]))
    photon_loss = photon_loss_all / m3                                       # evolve.F90:525
    st["total_photon_loss"] = photon_loss * dt * m3                          # :262-263
    st["totalsrc"] = float(np.sum(p["normflux"])) * p["S_star"] * dt         # :265
    st["photcons"] = (st["total_ion"] + 0.0 - st["totcollisions"]) / st["totalsrc"]   # :266
    return st


def evolve3d_py(p, dt, thick, thin, c):
    """evolve3D (evolve.F90:83-281), restart == 0, one rank: the outer iteration over pass_all_sources (:444-495, sources
    in file order) and global_pass (:499-573) with its two ways of ending, then the photon statistics of the step.
    Returns (niter, conv_flag per iteration, xh, xh_av, statistics)."""
    mesh = p["mesh"]
    ncell = mesh[0] * mesh[1] * mesh[2]
    xh_before = p["xh"].copy()
    xh = p["xh"].copy()
    xh_av, xh_intermed = xh.copy(), xh.copy()                       # :140-147
    niter = 0                                                       # :148
    conv_flag = ncell                                               # :149
    prev1 = prev0 = float(np.float32(2.0) * np.float32(mesh[0]) * np.float32(mesh[1]) * np.float32(mesh[2]))   # :150-151
    nsrc = len(p["normflux"])
    conv_criterion = min(int(c.convergence_fraction * mesh[0] * mesh[1] * mesh[2]), (nsrc - 1) // 3)   # :162
    flags = []
    q = dict(p)
    photon_loss_all = 0.0
    while True:
        s1 = float(np.sum(xh_intermed))                             # :183
        s0 = float(np.float32(ncell)) - s1                          # :184
        rel1 = abs(s1 - prev1) / s1 if s1 > 0.0 else 1.0            # :187-191
        rel0 = abs(s0 - prev0) / s0 if s0 > 0.0 else 1.0            # :192-196
        if conv_flag < conv_criterion or (rel1 < c.convergence_fraction and rel0 < c.convergence_fraction):   # :212-214
            xh = xh_intermed.copy()                                 # :215-217
            break
        if niter > 100:                                             # :228
            break
        prev1, prev0 = s1, s0                                       # :236-237
        niter += 1                                                  # :240
        phih = np.zeros_like(xh)                                    # set_rates_to_zero, :430-440
        photon_loss_all = 0.0
        q["xh"] = xh_av                                             # evolve0D reads xh_av (evolve_point.F90:137)
        for ns in range(1, nsrc + 1):                               # pass_all_sources / do_grid, one rank
            r = do_source_py(q, ns, thick, thin, c)
            phih += r[1]
            photon_loss_all = photon_loss_all + r[3]                # photon_loss(1)=photon_loss(1)+photon_loss_src
        conv_flag, xh_intermed, xh_av = global_pass_py(p, xh, xh_av, xh_intermed, phih, dt, c, p["temper"])   # :269
        flags.append(conv_flag)
    stats = photon_statistics_py(p, xh_before, xh, xh_av, photon_loss_all, dt, c)   # :277-279
    return niter, flags, xh, xh_av, stats


@pytest.mark.parametrize("clumping", ["scalar", "grid"])
def test_evolve3d_second_transcription(clumping):
    """a whole evolve3D step in plain Python against the C restatement: number of outer iterations, the convergence
    counter of every iteration, final fractions, the photon-conservation statistics"""
    from problems import make_problem, setup_oracle
    p = make_problem(9, nsrc=4, seed=12, state="random", use_LLS=True, flux=3e7, clumping=clumping)
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    thick, thin, _ = O.rad_ini()
    c = O.constants()
    dt = 1e6 * c.YEAR
    o = setup_oracle(p)
    ro = o.evolve3D(dt)
    niter, flags, xh, xh_av, stats = evolve3d_py(p, dt, thick, thin, c)
    assert niter == ro.niter and niter >= 2
    assert flags == list(ro.conv_flag[1:ro.niter + 1])
    np.testing.assert_allclose(o.xh, xh, rtol=0, atol=1e-12)
    np.testing.assert_allclose(o.xh_av, xh_av, rtol=0, atol=1e-12)
    for name, want in stats.items():
        tol = 1e-9 if name in ("dh0", "total_ion", "photcons") else 1e-11   # differences of nearly equal sums
        assert getattr(ro.final_stats, name) == pytest.approx(want, rel=tol), name
    assert ro.final_stats.LLS_loss == 0.0


def test_photo_tables_against_independent_quadrature():
    """stellar_photo_thick_table(tau) = S_star * int SED exp(-tau*s) dnu / int SED dnu and the thin table with one more
    factor s = (nu/nu_HI)^-2.8 (radiation_tables.F90:361-430, 524-543) for the 5e4 K black body.  The reference
    integrates 128 Romberg intervals over [nu_HI, 40 nu_HI]; against a fine trapezoid rule that coarse grid is good to
    0.5 % where the integrand peaks at the edge (small and moderate tau) and to 2e-5 at the largest tau -- the tolerance here is the
    reference's own quadrature error, which the restatement reproduces bit for bit (test_oracle.py, Appendix D)."""
    from c2ray3dm_b200 import constants as K
    thick, thin, d = O.rad_ini()
    assert d.freq_min == K.bb_MinFreq and d.freq_max == K.bb_MaxFreq
    nu = np.linspace(d.freq_min, d.freq_max, 400001)
    x = nu * K.hplanck / (K.k_B * K.bb_Teff)
    sed = nu * nu / np.expm1(x)
    cs = (nu / d.freq_min) ** (-K.pl_index_cross_section_HI)
    norm = np.trapezoid(sed, nu)
    for it, tol in ((0, 6e-3), (1, 6e-3), (1000, 6e-3), (1500, 6e-3), (1668, 6e-3), (1700, 6e-3), (1800, 3e-3), (1900, 1e-3),
                    (2000, 1e-4)):
        tau = 0.0 if it == 0 else 10.0 ** (-20.0 + 0.012 * (it - 1))
        assert thick[it] / thick[0] == pytest.approx(np.trapezoid(sed * np.exp(-tau * cs), nu) / norm, rel=tol)
        assert thin[it] / thick[0] == pytest.approx(np.trapezoid(sed * cs * np.exp(-tau * cs), nu) / norm, rel=tol)



def thermal_py(dt, T_initial, T_final, T_average, ndens_electron, ndens_atom, h_old1, h_av1, h1, heating, cool, c, zred,
               cosmological):
    """thermal (thermal.f90:22-176) with temper2pressr / pressr2temper (tped.f90:41-70), coolin (cooling.f90:38-59)
    and cosmo_cool (cosmology.F90:198-225).  Returns (final, average); both unchanged when T_initial <= minitemp."""
    mintemp, dtemp, cie = cool
    gamma1 = 5.0 / 3.0 - 1.0                                              # atomic.f90:23-25
    minitemp, relative_denergy = 1.0, float(np.float32(0.1))             # c2ray_parameters.f90:108,110
    el = lambda x1: ndens_atom * (x1 + c.abu_c)                           # electrondens, tped.f90:81
    internal_energy = (ndens_atom + el(h_old1)) * c.k_B * T_initial / gamma1      # thermal.f90:75-76
    if cosmological:                                                      # :81-85
        dzdt = c.H0 * (1. + zred) * math.sqrt(c.Omega0 * (1. + zred) ** 3 + 1. - c.Omega0)   # cosmology.F90:217
        cosmo_cool_rate = internal_energy * 2.0 / (1.0 + zred) * dzdt     # :220
    else:
        cosmo_cool_rate = 0.0
    if not T_initial > minitemp:                                          # :88
        return T_final, T_average
    cumulative_time = 0.0
    i_heating = 0
    T_average = 0.0
    T_mid = T_initial
    while True:
        i_heating += 1
        tpos = (math.log10(T_mid) - mintemp) / dtemp + 1.0                # coolin, cooling.f90:49-52
        itpos = min(61 - 1, max(1, int(tpos)))
        dtpos = tpos - float(itpos)
        itpos1 = min(61, itpos + 1)
        cooling = ndens_atom * ndens_electron * (cie[itpos - 1] + (cie[itpos1 - 1] - cie[itpos - 1]) * dtpos) \
            + cosmo_cool_rate                                             # thermal.f90:104-105
        thermal_rate = max(1e-50, abs(cooling - heating))                 # :108
        thermal_timescale = internal_energy / abs(thermal_rate)           # :109
        dt_thermal = relative_denergy * thermal_timescale                 # :112
        dt_ode = min(dt_thermal, dt - cumulative_time)                    # :113
        internal_energy = internal_energy + dt_ode * (heating - cooling)  # :116
        T_average = T_average + 0.5 * T_mid * dt_ode                      # :119-120
        T_mid = internal_energy * gamma1 / (c.k_B * (ndens_atom + el(h_av1)))     # :123-124
        T_average = T_average + 0.5 * T_mid * dt_ode                      # :125-126
        if T_mid < minitemp:                                              # :129-133 (no division by gamma1 there)
            internal_energy = (ndens_atom + el(h_av1)) * c.k_B * minitemp
            T_mid = minitemp
        cumulative_time = cumulative_time + dt_ode                        # :136
        if cumulative_time >= dt or abs(cumulative_time - dt) < float(np.float32(1e-6)) * dt:   # :141
            break
        if i_heating > 10000:                                             # :144
            break
    T_average = T_average / dt if dt > 0.0 else T_initial                 # :148-152
    T_final = internal_energy * gamma1 / (c.k_B * (ndens_atom + el(h1)))  # :155-156
    return T_final, T_average


def global_pass_thermal_py(p, xh, xh_av, xh_intermed, phih, phiheat, Tgrid, dt, c, cool, zred, cosmological):
    """evolve0D_global + do_chemistry (evolve_point.F90:305-555) with isothermal=.false.; Tgrid[...,0:3] = current,
    average, intermed as default reals (temperature_module.F90:21-31,134-169).  Updates Tgrid in place."""
    new_int, new_av = xh_intermed.copy(), xh_av.copy()
    conv_flag = 0
    eps, mfc, mfa = c.epsilon, c.minimum_fractional_change, c.minimum_fraction_of_atoms
    for idx in np.ndindex(xh.shape):
        h_old1 = max(eps, float(xh[idx]))
        h_av1 = max(eps, float(xh_av[idx]))
        h_old0, h_av0 = 1.0 - h_old1, 1.0 - h_av1
        ndens_p = float(p["ndens"][idx])
        T_start = [float(Tgrid[idx + (k,)]) for k in range(3)]            # get_temperature_point, :358 (current, average, intermed)
        ph, heat = float(phih[idx]), float(phiheat[idx])                  # :363-364
        clump = float(p["clumping_grid"][idx]) if p["clumping_grid"] is not None else float(np.float32(p["clumping"]))
        T_end_cur, T_end_avg, T_end_int = T_start                         # temperature_end=temperature_start, :437-438
        nit = 0
        while True:
            nit += 1
            yh0_av_old = h_av0
            de = ndens_p * (h_av1 + c.abu_c)
            (h0, h1), (h_av0, h_av1) = doric_py(dt, T_end_avg, de, (h_old0, h_old1), None, ph, clump, c)   # :513-514
            de = ndens_p * (h_av1 + c.abu_c)                              # :516
            T_end_int, T_end_avg = thermal_py(dt, T_start[0], T_end_int, T_end_avg, de, ndens_p, h_old1, h_av1, h1,
                                              heat, cool, c, zred, cosmological)   # :519-526
            if abs((h_av0 - yh0_av_old) / h_av0) < mfc or h_av0 < mfa:    # :530-535 (temperature_end%current never changes)
                break
            if nit > 400:
                break
        Tgrid[idx + (2,)] = np.float32(T_end_int)                         # set_temperature_point, :553
        Tgrid[idx + (1,)] = np.float32(T_end_avg)
        T_new_avg = float(Tgrid[idx + (1,)])                              # get_temperature_point, :380
        yh0_prev = 1.0 - max(eps, float(xh_av[idx]))
        if (abs(h_av0 - yh0_prev) > mfc and abs((h_av0 - yh0_prev) / h_av0) > mfc and h_av0 > mfa) or \
                (abs((T_start[1] - T_new_avg) / T_new_avg) > 1.0e-1 and abs(T_start[1] - T_new_avg) > 100.0):   # :382-388
            conv_flag += 1
        new_int[idx] = h1
        new_av[idx] = h_av1
    return conv_flag, new_int, new_av


def test_thermal_pass_second_transcription():
    """isothermal=.false.: one pass over the sources with heating rates, then two per-cell passes with the energy
    equation, against the C restatement (rates and heating to 1e-10, fractions 1e-12, temperatures as stored floats)"""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from problems import make_problem
    from thermal_common import cooling_table, setup_thermal_oracle
    tables4 = O.rad_ini_heat()
    p = make_problem(8, nsrc=3, seed=15, state="random", use_LLS=True, flux=3e8)
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    c = O.constants()
    zred = 9.0
    o = setup_thermal_oracle(p, tables4, zred=zred, cosmological=True)
    o.xh_av[...] = p["xh"]
    o.xh_intermed[...] = p["xh"]
    o.state_before()
    o.set_rates_to_zero()
    r = o.pass_all_sources()
    phih = np.zeros_like(p["xh"])
    phiheat = np.zeros_like(p["xh"])
    for ns in (1, 2, 3):
        res = do_source_py(p, ns, tables4[0], tables4[1], c, heat=(tables4[2], tables4[3]))
        phih += res[1]
        phiheat += res[5]
    np.testing.assert_allclose(o.phih, phih, rtol=1e-10, atol=0)
    assert np.array_equal(o.phiheat != 0, phiheat != 0)
    np.testing.assert_allclose(o.phiheat, phiheat, rtol=1e-10, atol=0)
    lt, lc = cooling_table()
    cool = (float(lt[0]), float(lt[1] - lt[0]), 10.0 ** lc)              # setup_cool, cooling.f90:77-85
    dt = 1e6 * c.YEAR
    Tgrid = o.temperature_grid.copy()
    for it in range(2):
        xh_av_in, xh_int_in = o.xh_av.copy(), o.xh_intermed.copy()
        conv, want_int, want_av = global_pass_thermal_py(p, o.xh.copy(), xh_av_in, xh_int_in, phih, phiheat, Tgrid, dt, c,
                                                         cool, zred, True)
        g = o.global_pass(dt, r.photon_loss_all)
        assert g.conv_flag == conv
        np.testing.assert_allclose(o.xh_intermed, want_int, rtol=0, atol=1e-12)
        np.testing.assert_allclose(o.xh_av, want_av, rtol=0, atol=1e-12)
        np.testing.assert_allclose(o.temperature_grid, Tgrid, rtol=3e-7, atol=0)   # stored as default reals
    assert float(Tgrid[..., 2].max()) > 1.01e4


def romberg_weights_py(nmax):
    """romberg_initialisation (romberg.f90:22-90): romw(0:nmax, pmax) for nmax = 2**pmax; the literals -1.0, 4.0 are
    default reals, so b(k) is formed in single precision before it is stored in real(dp)"""
    pmax = int(round(math.log(float(nmax)) / math.log(2.0)))
    f32 = np.float32
    a, b = [0.0] * (pmax + 1), [0.0] * (pmax + 1)
    for k in range(1, pmax + 1):
        b[k] = float(f32(-1.0) / (f32(4.0) ** f32(k) - f32(1.0)))     # :40 (4.0**k is exact in single precision here)
        a[k] = -b[k] * float(f32(4.0) ** f32(k))                      # :41
    s = [[0.0] * (pmax + 1) for _ in range(pmax + 1)]
    romw = [[0.0] * (pmax + 1) for _ in range(nmax + 1)]              # romw[j][i]
    for k in range(0, pmax + 1):                                      # :53-66
        s[k][0] = 1.0
        for j in range(1, pmax + 1):
            for i in range(pmax, j - 1, -1):
                s[i][j] = a[j] * s[i][j - 1] + b[j] * s[i - 1][j - 1]
        for i in range(k, pmax + 1):
            for j in range(0, 2 ** k + 1):
                romw[2 ** (i - k) * j][i] = s[i][i] * 2 ** (i - k) + romw[2 ** (i - k) * j][i]
        s[k][0] = 0.0
    for i in range(0, pmax + 1):                                      # :70-73
        romw[0][i] = 0.5 * romw[0][i]
        romw[2 ** i][i] = 0.5 * romw[2 ** i][i]
    return [romw[j][pmax] for j in range(nmax + 1)]


def test_romberg_weights_and_tables_second_transcription():
    """the Romberg weights bit for bit, and the photo and heat tables re-integrated in Python from the restatement's SED
    normalisation: spec_integration / fill_photo_integrands / fill_heating_integrands / make_photo_tables
    (radiation_tables.F90:130-236, 361-543) with BB_SED (:434-452) and Vector_Romberg (romberg.f90:158-187)"""
    from c2ray3dm_b200 import constants as K
    thick, thin, d = O.rad_ini()
    hthick, hthin = O.rad_ini_heat()[2:]
    c = O.constants()
    w = romberg_weights_py(128)
    assert w == list(d.romw7)
    nf = 128
    freq = [d.freq_min + d.delta_freq * float(np.float32(i)) for i in range(nf + 1)]       # :264-272
    cs = [(f / d.freq_min) ** (-K.pl_index_cross_section_HI) for f in freq]                  # :276-297
    r2 = d.R_star * d.R_star
    sed = [4.0 * c.pi * r2 * c.two_pi_over_c_square * f * f / (math.exp(f * d.h_over_kT) - 1.0)
           if f * d.h_over_kT < 700.0 else 0.0 for f in freq]                                # BB_SED
    for it in (0, 1, 500, 1000, 1500, 1668, 1700, 1800, 1900, 2000):
        tau = 0.0 if it == 0 else float(np.float32(10.0)) ** (c.minlogtau + c.dlogtau * float(np.float32(it - 1)))   # :250-256
        a_thick = a_thin = h_thick = h_thin = 0.0
        for i in range(nf + 1):
            if tau * cs[i] < 700.0:
                f_thick = sed[i] * math.exp(-tau * cs[i])
                f_thin = sed[i] * cs[i] * math.exp(-tau * cs[i])
            else:
                f_thick = f_thin = 0.0
            a_thick = a_thick + f_thick * d.delta_freq * w[i]
            a_thin = a_thin + f_thin * d.delta_freq * w[i]
            h_thick = h_thick + (c.hplanck * (freq[i] - c.ion_freq_HI) * f_thick) * d.delta_freq * w[i]   # :482-487
            h_thin = h_thin + (c.hplanck * (freq[i] - c.ion_freq_HI) * f_thin) * d.delta_freq * w[i]
        assert thick[it] == pytest.approx(a_thick, rel=1e-13)
        assert thin[it] == pytest.approx(a_thin, rel=1e-13)
        assert hthick[it] == pytest.approx(h_thick, rel=1e-13)
        assert hthin[it] == pytest.approx(h_thin, rel=1e-13)
