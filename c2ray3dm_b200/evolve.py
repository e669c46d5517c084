"""Host-side mirror of the reference's `module evolve` and of the module state it reads.

The reference exposes one entry point, `evolve3D(time,dt,restart)` (evolve.F90:83), and takes all
other inputs from module variables (`ndens`, `xh`, `dr`, `vol`, `clumping`, `coldensh_LLS`,
`srcpos`, `NormFlux_stellar`, the photo-ionization tables ...).  `Evolve` keeps the same names;
every setter forwards to the C ABI, which owns the device copies.  Grids are numpy arrays in
Fortran element order (i fastest), i.e. C-order shape (mesh3, mesh2, mesh1) or flat.
"""
import ctypes as C

import numpy as np

from . import constants as K
from . import lib as _lib


class C2RayError(RuntimeError):
    pass


def shard_sources(NumSrc, rank, npr):
    """1-based source numbers traced by `rank`: do ns1=1+rank,NumSrc,npr (master_slave.F90:85)."""
    return list(range(1 + rank, NumSrc + 1, npr))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Evolve:
    def __init__(self, mesh, device=0, rank=0, nranks=1, type_of_clumping=1, use_LLS=True,
                 type_of_LLS=1, **overrides):
        self.h = None
        self.L = _lib.load()
        if np.isscalar(mesh):
            mesh = (int(mesh),) * 3
        self.mesh = tuple(int(m) for m in mesh)
        cfg = _lib.default_config()
        cfg.mesh[0], cfg.mesh[1], cfg.mesh[2] = self.mesh
        cfg.device = device
        cfg.rank = rank
        cfg.nranks = nranks
        cfg.type_of_clumping = type_of_clumping
        cfg.use_LLS = int(bool(use_LLS))
        cfg.type_of_LLS = type_of_LLS
        for k, v in overrides.items():
            if not hasattr(cfg, k):
                raise C2RayError("unknown configuration field %r" % k)
            setattr(cfg, k, v)
        self.cfg = cfg
        self.ncell = self.mesh[0] * self.mesh[1] * self.mesh[2]
        self.shape = (self.mesh[2], self.mesh[1], self.mesh[0])
        h = C.c_void_p()
        rc = self.L.c2b_create(C.byref(cfg), C.byref(h))
        if rc:
            raise C2RayError("c2b_create failed (%d): %s" % (rc, self.L.c2b_last_error(None).decode()))
        self.h = h
        self.NumSrc = 0
        self.last_report = None

    def close(self):
        if getattr(self, "h", None):
            self.L.c2b_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc:
            raise C2RayError("%s failed (%d): %s" % (what, rc, self.L.c2b_last_error(self.h).decode()))

    # ---- multi-GPU ---------------------------------------------------------------------------
    @staticmethod
    def get_unique_id():
        buf = C.create_string_buffer(_lib.UNIQUE_ID_BYTES)
        L = _lib.load()
        if L.c2b_get_unique_id(buf):
            raise C2RayError("c2b_get_unique_id failed: %s" % L.c2b_last_error(None).decode())
        return buf.raw

    def comm_init(self, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), _lib.UNIQUE_ID_BYTES)
        self._ck(self.L.c2b_comm_init(self.h, buf), "c2b_comm_init")

    # ---- module state -------------------------------------------------------------------------
    def rad_ini(self, T_eff=K.bb_Teff, S_star=K.bb_S_star, freq_min=K.bb_MinFreq, freq_max=K.bb_MaxFreq):
        """rad_ini (radiation_tables.F90:95-126) for the black-body SED; returns (thick, thin)."""
        thick = np.zeros(_lib.NUMTAU + 1)
        thin = np.zeros(_lib.NUMTAU + 1)
        self._ck(self.L.c2b_rad_ini_blackbody(self.h, T_eff, S_star, freq_min, freq_max, K.hplanck, K.k_B,
                                              K.two_pi_over_c_square, K.R_SOLAR,
                                              K.pl_index_cross_section_HI, _dp(thick), _dp(thin)),
                 "c2b_rad_ini_blackbody")
        return thick, thin

    def set_tables(self, thick, thin):
        thick = np.ascontiguousarray(thick, dtype=np.float64)
        thin = np.ascontiguousarray(thin, dtype=np.float64)
        self._ck(self.L.c2b_set_tables(self.h, _dp(thick), _dp(thin), thick.size), "c2b_set_tables")

    # ---- non-isothermal inputs (Evolve(..., isothermal=0)) -------------------------------------------
    def set_heat_tables(self, heat_thick, heat_thin):
        a, b = (np.ascontiguousarray(x, dtype=np.float64) for x in (heat_thick, heat_thin))
        self._ck(self.L.c2b_set_heat_tables(self.h, _dp(a), _dp(b), a.size), "c2b_set_heat_tables")

    def heat_tables(self):
        """stellar_heat_thick_table / ..thin.. as the device holds them (rad_ini builds them when not isothermal)"""
        a, b = np.zeros(_lib.NUMTAU + 1), np.zeros(_lib.NUMTAU + 1)
        self._ck(self.L.c2b_get_heat_tables(self.h, _dp(a), _dp(b)), "c2b_get_heat_tables")
        return a, b

    def set_cooling_table(self, log10_temp, log10_cool):
        """the 61 rows of tables/corocool.tab (cooling.f90:64-87)"""
        a, b = (np.ascontiguousarray(x, dtype=np.float64) for x in (log10_temp, log10_cool))
        self._ck(self.L.c2b_set_cooling_table(self.h, _dp(a), _dp(b), a.size), "c2b_set_cooling_table")

    def set_redshift(self, zred):
        self._ck(self.L.c2b_set_redshift(self.h, float(zred)), "c2b_set_redshift")

    def set_temperature_grid(self, tg):
        a = np.ascontiguousarray(tg, dtype=np.float32).reshape(-1)
        if a.size != 3 * self.ncell:
            raise C2RayError("temperature_grid size mismatch (3 values per cell)")
        self._ck(self.L.c2b_set_temperature_grid(self.h, _fp(a)), "c2b_set_temperature_grid")

    @property
    def temperature_grid(self):
        """(n3, n2, n1, 3) float32: current, average, intermed (temperature_module.F90:21-25)"""
        out = np.empty(3 * self.ncell, dtype=np.float32)
        self._ck(self.L.c2b_get_temperature_grid(self.h, _fp(out)), "c2b_get_temperature_grid")
        return out.reshape(self.shape + (3,))

    @property
    def phiheat_grid(self):
        return self._get("c2b_get_phiheat")

    def set_density(self, ndens):
        a = np.ascontiguousarray(ndens, dtype=np.float32).reshape(-1)
        if a.size != self.ncell:
            raise C2RayError("ndens has %d elements, mesh has %d" % (a.size, self.ncell))
        self._ck(self.L.c2b_set_density(self.h, _fp(a)), "c2b_set_density")

    def set_geometry(self, dr, vol=None):
        dr = np.ascontiguousarray(np.broadcast_to(np.asarray(dr, dtype=np.float64), (3,)))
        if vol is None:
            vol = dr[0] * dr[1] * dr[2]  # grid.F90:131
        self.dr, self.vol = dr.copy(), float(vol)
        self._ck(self.L.c2b_set_geometry(self.h, _dp(dr), float(vol)), "c2b_set_geometry")

    def cosmo_evol(self, zfactor):
        """cosmology.F90:161-193 applied to the device copies of dr, vol and ndens."""
        self._ck(self.L.c2b_cosmo_evol(self.h, float(zfactor)), "c2b_cosmo_evol")

    def set_clumping(self, clumping):
        """scalar (type_of_clumping 1,2) or float32 grid (3,4,5); clumping_module.F90:17-18"""
        if np.isscalar(clumping):
            self._ck(self.L.c2b_set_clumping_scalar(self.h, float(clumping)), "c2b_set_clumping_scalar")
        else:
            a = np.ascontiguousarray(clumping, dtype=np.float32).reshape(-1)
            if a.size != self.ncell:
                raise C2RayError("clumping grid size mismatch")
            self._ck(self.L.c2b_set_clumping_grid(self.h, _fp(a)), "c2b_set_clumping_grid")

    def set_LLS(self, coldensh_LLS=None, LLS_grid=None, R_max_LLS=None):
        if coldensh_LLS is not None:
            self._ck(self.L.c2b_set_lls_scalar(self.h, float(coldensh_LLS)), "c2b_set_lls_scalar")
        if LLS_grid is not None:
            a = np.ascontiguousarray(LLS_grid, dtype=np.float32).reshape(-1)
            if a.size != self.ncell:
                raise C2RayError("LLS grid size mismatch")
            self._ck(self.L.c2b_set_lls_grid(self.h, _fp(a)), "c2b_set_lls_grid")
        if R_max_LLS is not None:
            self._ck(self.L.c2b_set_lls_rmax(self.h, float(R_max_LLS)), "c2b_set_lls_rmax")

    def set_temperature(self, temper_val):
        self._ck(self.L.c2b_set_temperature(self.h, float(temper_val)), "c2b_set_temperature")

    def set_sources(self, srcpos, NormFlux_stellar, S_star=K.bb_S_star):
        """srcpos: (NumSrc,3) 1-based mesh positions; NormFlux_stellar: (NumSrc,)"""
        srcpos = np.ascontiguousarray(srcpos, dtype=np.int32).reshape(-1, 3)
        nf = np.ascontiguousarray(NormFlux_stellar, dtype=np.float64).reshape(-1)
        if srcpos.shape[0] != nf.size:
            raise C2RayError("srcpos and NormFlux_stellar disagree on NumSrc")
        self.NumSrc = int(nf.size)
        self._ck(self.L.c2b_set_sources(self.h, self.NumSrc, srcpos.ctypes.data_as(C.POINTER(C.c_int32)),
                                        _dp(nf), float(S_star)), "c2b_set_sources")

    def set_xh(self, xh):
        a = np.ascontiguousarray(xh, dtype=np.float64).reshape(-1)
        if a.size != self.ncell:
            raise C2RayError("xh size mismatch")
        self._ck(self.L.c2b_set_xh(self.h, _dp(a)), "c2b_set_xh")

    # ---- the hot path --------------------------------------------------------------------------
    def evolve3D(self, time, dt, restart=0):
        """evolve3D(time,dt,restart), evolve.F90:83.  Returns the step report (what the reference logs)."""
        rep = _lib.StepReport()
        self._ck(self.L.c2b_evolve3d(self.h, float(time), float(dt), int(restart), C.byref(rep)), "c2b_evolve3d")
        self.last_report = rep
        return rep

    def begin_step(self):
        s = C.c_double()
        self._ck(self.L.c2b_begin_step(self.h, C.byref(s)), "c2b_begin_step")
        return s.value

    def pass_all_sources(self, niter=1, dt=0.0):
        rep = _lib.PassReport()
        self._ck(self.L.c2b_pass_all_sources(self.h, niter, dt, C.byref(rep)), "c2b_pass_all_sources")
        return rep

    def global_pass(self, dt):
        rep = _lib.GlobalReport()
        self._ck(self.L.c2b_global_pass(self.h, float(dt), C.byref(rep)), "c2b_global_pass")
        return rep

    def end_step(self, dt, converged=True):
        st = _lib.PhotonStats()
        self._ck(self.L.c2b_end_step(self.h, float(dt), int(bool(converged)), C.byref(st)), "c2b_end_step")
        return st

    # ---- outputs --------------------------------------------------------------------------------
    def _get(self, fn, dtype=np.float64):
        out = np.empty(self.ncell, dtype=dtype)
        ptr = _dp(out) if dtype == np.float64 else _fp(out)
        self._ck(getattr(self.L, fn)(self.h, ptr), fn)
        return out.reshape(self.shape)

    @property
    def xh(self):
        return self._get("c2b_get_xh")

    @property
    def xh_av(self):
        return self._get("c2b_get_xh_av")

    @property
    def xh_intermed(self):
        return self._get("c2b_get_xh_intermed")

    @property
    def phih_grid(self):
        return self._get("c2b_get_phih")

    @property
    def phih_grid_si(self):
        """real(phih_grid,si) as written to IonRates3D_*.bin (output.F90:359)"""
        return self._get("c2b_get_phih_f32", np.float32)

    def source_nbox(self):
        out = np.zeros(max(self.NumSrc, 1), dtype=np.int32)
        self._ck(self.L.c2b_get_source_nbox(self.h, out.ctypes.data_as(C.POINTER(C.c_int32))), "c2b_get_source_nbox")
        return out[:self.NumSrc]

    def source_loss(self):
        out = np.zeros(max(self.NumSrc, 1), dtype=np.float64)
        self._ck(self.L.c2b_get_source_loss(self.h, _dp(out)), "c2b_get_source_loss")
        return out[:self.NumSrc]

    def get_iter_state(self):
        niter = C.c_int32()
        pl = C.c_double()
        phih = np.empty(self.ncell)
        xav = np.empty(self.ncell)
        xint = np.empty(self.ncell)
        self._ck(self.L.c2b_get_iter_state(self.h, C.byref(niter), C.byref(pl), _dp(phih), _dp(xav), _dp(xint)),
                 "c2b_get_iter_state")
        return niter.value, pl.value, phih.reshape(self.shape), xav.reshape(self.shape), xint.reshape(self.shape)

    def set_iter_state(self, niter, photon_loss_all, phih, xh_av, xh_intermed):
        a = [np.ascontiguousarray(x, dtype=np.float64).reshape(-1) for x in (phih, xh_av, xh_intermed)]
        self._ck(self.L.c2b_set_iter_state(self.h, int(niter), float(photon_loss_all), _dp(a[0]), _dp(a[1]), _dp(a[2])),
                 "c2b_set_iter_state")

    def trace_source_debug(self, ns):
        """do_source(ns) alone: returns (coldensh_out, phih, nbox, photon_loss_src)."""
        cd = np.empty(self.ncell)
        ph = np.empty(self.ncell)
        nbox = C.c_int32()
        loss = C.c_double()
        self._ck(self.L.c2b_trace_source_debug(self.h, int(ns), _dp(cd), _dp(ph), C.byref(nbox), C.byref(loss)),
                 "c2b_trace_source_debug")
        return cd.reshape(self.shape), ph.reshape(self.shape), nbox.value, loss.value

    def measure_dfma_rate(self):
        r = C.c_double()
        self._ck(self.L.c2b_measure_dfma_rate(self.h, C.byref(r)), "c2b_measure_dfma_rate")
        return r.value

    def set_clumping_from_density(self, p1, p2, p3, avg_dens):
        """deterministic_clumping (clumping_module.F90:327-363) on the device-resident density"""
        self._ck(self.L.c2b_set_clumping_from_density(self.h, float(p1), float(p2), float(p3), float(avg_dens)),
                 "c2b_set_clumping_from_density")

    @property
    def clumping_grid(self):
        out = np.empty(self.ncell, dtype=np.float32)
        self._ck(self.L.c2b_get_clumping_grid(self.h, _fp(out)), "c2b_get_clumping_grid")
        return out.reshape(self.shape)

    def source_owner(self):
        """rank that traces each source in the next pass (multi-rank load balance)"""
        a = np.zeros(max(1, self.NumSrc), dtype=np.int32)
        self._ck(self.L.c2b_get_source_owner(self.h, a.ctypes.data_as(C.POINTER(C.c_int32))), "c2b_get_source_owner")
        return a[:self.NumSrc]

    def route_counts(self):
        """sources dealt so far to (one CTA, one cluster, one warp for the first subbox, handed over by the warp shape)"""
        a = (C.c_int64 * 4)()
        self._ck(self.L.c2b_get_route_counts(self.h, a), "c2b_get_route_counts")
        return tuple(int(x) for x in a)

    def synchronize(self):
        self._ck(self.L.c2b_synchronize(self.h), "c2b_synchronize")

    def save_xh(self):
        """keeps a device copy of xh (harness: a benchmark restarts every step from it)"""
        self._ck(self.L.c2b_save_xh_dev(self.h), "c2b_save_xh_dev")

    def restore_xh(self):
        self._ck(self.L.c2b_restore_xh_dev(self.h), "c2b_restore_xh_dev")
