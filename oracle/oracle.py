"""ctypes binding of the CPU oracle (oracle/c2ray_oracle.c).

TEST INFRASTRUCTURE ONLY: may be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from the product package c2ray3dm_b200.

PARITY UNPINNED by the reference itself (no tests / golden vectors there, no Fortran compiler
here); see DESIGN.md for the pins this repository adds.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libc2ray_oracle.so")
NUMTAU = 2000
NUMFREQ = 128
MAX_ITER = 104


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("c2ray_oracle.c", "c2ray_oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


class Constants(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "pi sigma_HI_at_ion_freq eth0 ev2k temph0 colh0 ev2fr ion_freq_HI ion_freq_HeII "
        "bb_MaxFreq two_pi_over_c_square bh00 albpow hplanck k_B c_light m_p sigma_SB "
        "abu_he abu_c mu h Omega0 Omega_B Mpc H0 rho_crit_0 YEAR R_SOLAR "
        "epsilon convergence_fraction minimum_fractional_change minimum_fraction_of_atoms "
        "loss_fraction max_coldensh tau_photo_limit sqrt3 sqrt2 minlogtau dlogtau "
        "xh_initial bb_Teff bb_S_star").split()]


class RadDiag(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                "S_star_unscaled S_scaling R_star h_over_kT freq_min freq_max delta_freq".split()] + \
               [("romw7", C.c_double * (NUMFREQ + 1))]


class SourceReport(C.Structure):
    _fields_ = [("nbox", C.c_int), ("photon_loss_src", C.c_double), ("updates", C.c_int64)]


class PassReport(C.Structure):
    _fields_ = [("photon_loss_all", C.c_double), ("sum_nbox_all", C.c_int64), ("updates", C.c_int64)]


class PhotonStats(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                "h0_before h1_before h0_after h1_after totrec totcollisions dh0 total_ion "
                "totalsrc photcons total_photon_loss LLS_loss".split()]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class GlobalReport(C.Structure):
    _fields_ = [("conv_flag", C.c_int), ("min_avg_neutral", C.c_double),
                ("sum_xh_intermed", C.c_double), ("stats", PhotonStats)]


class StepReport(C.Structure):
    _fields_ = [("niter", C.c_int), ("converged", C.c_int), ("conv_criterion", C.c_int),
                ("conv_flag", C.c_int * MAX_ITER),
                ("rel_change_sum_xh1", C.c_double * MAX_ITER),
                ("rel_change_sum_xh0", C.c_double * MAX_ITER),
                ("photon_loss_all", C.c_double * MAX_ITER),
                ("sum_nbox_all", C.c_int64 * MAX_ITER),
                ("updates", C.c_int64 * MAX_ITER),
                ("iter_stats", PhotonStats * MAX_ITER),
                ("final_stats", PhotonStats),
                ("grtotal_ion", C.c_double), ("grtotal_src", C.c_double),
                ("total_updates", C.c_int64),
                ("seconds_raytrace", C.c_double), ("seconds_global", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        vp, dp, fp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.c_int] * 3
        L.orc_destroy.argtypes = [vp]
        L.orc_get_constants.argtypes = [C.POINTER(Constants)]
        L.orc_rad_ini.argtypes = [dp, dp, C.POINTER(RadDiag)]
        L.orc_set_tables.argtypes = [vp, dp, dp]
        L.orc_set_density.argtypes = [vp, fp]
        L.orc_set_geometry.argtypes = [vp, dp, C.c_double]
        L.orc_set_clumping.argtypes = [vp, C.c_int, C.c_float, fp]
        L.orc_deterministic_clumping.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_clumping_grid.argtypes = [vp]
        L.orc_clumping_grid.restype = fp
        L.orc_set_lls.argtypes = [vp, C.c_int, C.c_int, C.c_double, fp, C.c_double]
        L.orc_set_temperature.argtypes = [vp, C.c_double]
        L.orc_set_sources.argtypes = [vp, C.c_int, ip, dp, C.c_double]
        L.orc_set_xh.argtypes = [vp, dp]
        L.orc_set_xh_av.argtypes = [vp, dp]
        L.orc_set_loss_fraction.argtypes = [vp, C.c_double]
        L.orc_set_walk_order.argtypes = [vp, C.c_int]
        L.orc_set_rank.argtypes = [vp, C.c_int, C.c_int]
        L.orc_set_threads.argtypes = [vp, C.c_int]
        L.orc_set_omp_in_source.argtypes = [vp, C.c_int]
        for f in ("orc_xh", "orc_xh_av", "orc_xh_intermed", "orc_phih", "orc_coldensh_out"):
            getattr(L, f).restype = dp
            getattr(L, f).argtypes = [vp]
        L.orc_cinterp.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), dp, dp]
        L.orc_photoion_rates.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_double, dp]
        L.orc_doric.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_double, dp, dp,
                                C.c_double, C.c_float]
        L.orc_do_source.argtypes = [vp, C.c_int, C.POINTER(SourceReport)]
        L.orc_set_rates_to_zero.argtypes = [vp]
        L.orc_pass_all_sources.argtypes = [vp, C.POINTER(PassReport)]
        L.orc_global_pass.argtypes = [vp, C.c_double, C.c_double, C.POINTER(GlobalReport)]
        L.orc_state_before.argtypes = [vp]
        L.orc_calculate_photon_statistics.argtypes = [vp, C.c_double, dp, dp, C.POINTER(PhotonStats)]
        L.orc_evolve3D.argtypes = [vp, C.c_double, C.c_int, C.POINTER(StepReport)]
        L.orc_rad_ini_heat.argtypes = [dp, dp, dp, dp]
        L.orc_set_isothermal.argtypes = [vp, C.c_int]
        L.orc_set_heat_tables.argtypes = [vp, dp, dp]
        L.orc_set_cooling_table.argtypes = [vp, dp, dp]
        L.orc_set_redshift.argtypes = [vp, C.c_double, C.c_int]
        L.orc_temperature_grid.restype = fp
        L.orc_temperature_grid.argtypes = [vp]
        L.orc_phiheat.restype = dp
        L.orc_phiheat.argtypes = [vp]
        L.orc_set_dump_iteration.argtypes = [vp, C.c_int]
        L.orc_get_dump.argtypes = [vp, C.POINTER(C.c_int), dp, dp, dp, dp]
        L.orc_evolve3D_restart.argtypes = [vp, C.c_double, C.c_int, C.c_int, C.c_double, dp, dp, dp,
                                           C.POINTER(StepReport)]
        _lib = L
    return _lib


def constants():
    c = Constants()
    lib().orc_get_constants(C.byref(c))
    return c


_tables_cache = None


def rad_ini():
    """Returns (thick, thin, diag) as built by the restated rad_ini."""
    global _tables_cache
    if _tables_cache is None:
        thick = np.zeros(NUMTAU + 1)
        thin = np.zeros(NUMTAU + 1)
        d = RadDiag()
        lib().orc_rad_ini(_dp(thick), _dp(thin), C.byref(d))
        _tables_cache = (thick, thin, d)
    return _tables_cache


def rad_ini_heat():
    """rad_ini with isothermal=.false.: (thick, thin, heat_thick, heat_thin)"""
    t = [np.zeros(NUMTAU + 1) for _ in range(4)]
    lib().orc_rad_ini_heat(_dp(t[0]), _dp(t[1]), _dp(t[2]), _dp(t[3]))
    return tuple(t)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Oracle:
    """Module state of the reference + the hot-path entry points, on the CPU.

    Grids are numpy arrays in Fortran (i fastest) linear order, i.e. shape (m3, m2, m1) in C order.
    """

    def __init__(self, mesh):
        if np.isscalar(mesh):
            mesh = (int(mesh),) * 3
        self.mesh = tuple(int(m) for m in mesh)
        self.L = lib()
        self.h = self.L.orc_create(*self.mesh)
        self.ncell = self.mesh[0] * self.mesh[1] * self.mesh[2]
        self.shape = (self.mesh[2], self.mesh[1], self.mesh[0])
        thick, thin, _ = rad_ini()
        self.L.orc_set_tables(self.h, _dp(thick), _dp(thin))

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    # -- setters ------------------------------------------------------------------------------
    def set_tables(self, thick, thin):
        thick = np.ascontiguousarray(thick, dtype=np.float64)
        thin = np.ascontiguousarray(thin, dtype=np.float64)
        self.L.orc_set_tables(self.h, _dp(thick), _dp(thin))

    def set_density(self, ndens):
        a = np.ascontiguousarray(ndens, dtype=np.float32).reshape(-1)
        assert a.size == self.ncell
        self.L.orc_set_density(self.h, _fp(a))

    def set_geometry(self, dr, vol=None):
        dr = np.ascontiguousarray(np.broadcast_to(np.asarray(dr, dtype=np.float64), (3,)))
        if vol is None:
            vol = dr[0] * dr[1] * dr[2]
        self.L.orc_set_geometry(self.h, _dp(dr), float(vol))

    def set_clumping(self, type_of_clumping=1, clumping=1.0, grid=None):
        g = None
        if grid is not None:
            g = np.ascontiguousarray(grid, dtype=np.float32).reshape(-1)
        self.L.orc_set_clumping(self.h, int(type_of_clumping), float(clumping), _fp(g) if g is not None else None)

    def deterministic_clumping(self, p1, p2, p3, avg_dens):
        """deterministic_clumping (clumping_module.F90:327-363) from the density the oracle holds"""
        self.L.orc_deterministic_clumping(self.h, float(p1), float(p2), float(p3), float(avg_dens))

    @property
    def clumping_grid(self):
        return self._grid("orc_clumping_grid")

    def set_lls(self, use_LLS=False, type_of_LLS=1, coldensh_LLS=0.0, grid=None, R_max_LLS=0.0):
        g = None
        if grid is not None:
            g = np.ascontiguousarray(grid, dtype=np.float32).reshape(-1)
        self.L.orc_set_lls(self.h, int(bool(use_LLS)), int(type_of_LLS), float(coldensh_LLS),
                           _fp(g) if g is not None else None, float(R_max_LLS))

    def set_temperature(self, t):
        self.L.orc_set_temperature(self.h, float(t))

    def set_sources(self, srcpos, normflux, S_star=1e48):
        srcpos = np.ascontiguousarray(srcpos, dtype=np.int32).reshape(-1, 3)
        nf = np.ascontiguousarray(normflux, dtype=np.float64).reshape(-1)
        assert srcpos.shape[0] == nf.size
        self.L.orc_set_sources(self.h, int(nf.size), srcpos.ctypes.data_as(C.POINTER(C.c_int32)), _dp(nf), float(S_star))

    def set_xh(self, xh):
        a = np.ascontiguousarray(xh, dtype=np.float64).reshape(-1)
        assert a.size == self.ncell
        self.L.orc_set_xh(self.h, _dp(a))

    def set_xh_av(self, x):
        a = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        assert a.size == self.ncell
        self.L.orc_set_xh_av(self.h, _dp(a))

    def set_loss_fraction(self, lf):
        self.L.orc_set_loss_fraction(self.h, float(lf))

    def set_walk_order(self, order):
        self.L.orc_set_walk_order(self.h, int(order))

    def set_rank(self, rank, npr):
        self.L.orc_set_rank(self.h, int(rank), int(npr))

    def set_threads(self, n):
        self.L.orc_set_threads(self.h, int(n))

    # -- grids --------------------------------------------------------------------------------
    def set_omp_in_source(self, on):
        """threads inside one source (the OpenMP build, evolve_source.F90:141-186) instead of one source per thread"""
        self.L.orc_set_omp_in_source(self.h, int(bool(on)))

    def _grid(self, fn):
        p = getattr(self.L, fn)(self.h)
        return np.ctypeslib.as_array(p, shape=(self.ncell,)).reshape(self.shape)

    @property
    def xh(self):
        return self._grid("orc_xh")

    @property
    def xh_av(self):
        return self._grid("orc_xh_av")

    @property
    def xh_intermed(self):
        return self._grid("orc_xh_intermed")

    @property
    def phih(self):
        return self._grid("orc_phih")

    @property
    def coldensh_out(self):
        return self._grid("orc_coldensh_out")

    # -- hot path ------------------------------------------------------------------------------
    def cinterp(self, pos, srcpos):
        p = (C.c_int * 3)(*[int(v) for v in pos])
        s = (C.c_int * 3)(*[int(v) for v in srcpos])
        cd, path = C.c_double(), C.c_double()
        self.L.orc_cinterp(self.h, p, s, C.byref(cd), C.byref(path))
        return cd.value, path.value

    def photoion_rates(self, colum_in, colum_out, vol, normflux):
        out = np.zeros(3)
        self.L.orc_photoion_rates(self.h, colum_in, colum_out, vol, normflux, _dp(out))
        return tuple(out)

    def doric(self, dt, temp0, rhe, rhh, xfh, xfh_av, phih, clumping=1.0):
        a = np.array(xfh, dtype=np.float64)
        b = np.array(xfh_av, dtype=np.float64)
        self.L.orc_doric(self.h, dt, temp0, rhe, rhh, _dp(a), _dp(b), phih, clumping)
        return a, b

    def do_source(self, ns):
        r = SourceReport()
        self.L.orc_do_source(self.h, int(ns), C.byref(r))
        return r

    def set_rates_to_zero(self):
        self.L.orc_set_rates_to_zero(self.h)

    def pass_all_sources(self):
        r = PassReport()
        self.L.orc_pass_all_sources(self.h, C.byref(r))
        return r

    def global_pass(self, dt, photon_loss_all=0.0):
        r = GlobalReport()
        self.L.orc_global_pass(self.h, float(dt), float(photon_loss_all), C.byref(r))
        return r

    def state_before(self):
        self.L.orc_state_before(self.h)

    def evolve3D(self, dt, max_outer_iter=0):
        r = StepReport()
        self.L.orc_evolve3D(self.h, float(dt), int(max_outer_iter), C.byref(r))
        return r

    # ---- non-isothermal path ------------------------------------------------------------------------
    def set_isothermal(self, isothermal):
        """isothermal=.false. allocates phiheat_grid and temperature_grid (filled with temper_val)"""
        self.L.orc_set_isothermal(self.h, int(bool(isothermal)))

    def set_heat_tables(self, heat_thick, heat_thin):
        a, b = (np.ascontiguousarray(x, dtype=np.float64) for x in (heat_thick, heat_thin))
        self.L.orc_set_heat_tables(self.h, _dp(a), _dp(b))

    def set_cooling_table(self, log10_temp, log10_cool):
        """the 61 rows of tables/corocool.tab (cooling.f90:64-87)"""
        a, b = (np.ascontiguousarray(x, dtype=np.float64) for x in (log10_temp, log10_cool))
        assert a.size == 61 and b.size == 61
        self.L.orc_set_cooling_table(self.h, _dp(a), _dp(b))

    def set_redshift(self, zred, cosmological=True):
        self.L.orc_set_redshift(self.h, float(zred), int(bool(cosmological)))

    @property
    def temperature_grid(self):
        """(n3, n2, n1, 3) float32 view: current, average, intermed (temperature_module.F90:21-25)"""
        p = self.L.orc_temperature_grid(self.h)
        return np.ctypeslib.as_array(p, shape=(3 * self.ncell,)).reshape(self.shape + (3,))

    @property
    def phiheat(self):
        p = self.L.orc_phiheat(self.h)
        return np.ctypeslib.as_array(p, shape=(self.ncell,)).reshape(self.shape)

    def set_dump_iteration(self, niter):
        """write_iteration_dump (evolve.F90:285-324) after pass_all_sources of iteration `niter` (0 = never)"""
        self.L.orc_set_dump_iteration(self.h, int(niter))

    def get_dump(self):
        """(niter, photon_loss_all, phih_grid, xh_av, xh_intermed) of the last dump, or None"""
        n = C.c_int()
        pl = C.c_double()
        g = [np.empty(self.shape, dtype=np.float64) for _ in range(3)]
        if self.L.orc_get_dump(self.h, C.byref(n), C.byref(pl), _dp(g[0]), _dp(g[1]), _dp(g[2])):
            return None
        return n.value, pl.value, g[0], g[1], g[2]

    def evolve3D_restart(self, dt, niter, photon_loss_all, phih, xh_av, xh_intermed, max_outer_iter=0):
        """evolve3D(time,dt,restart/=0) with the dump record in place of iterdump[12].bin (evolve.F90:154-158)"""
        r = StepReport()
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (phih, xh_av, xh_intermed)]
        self.L.orc_evolve3D_restart(self.h, float(dt), int(max_outer_iter), int(niter), float(photon_loss_all),
                                    _dp(a[0]), _dp(a[1]), _dp(a[2]), C.byref(r))
        return r
