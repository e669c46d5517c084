"""ctypes binding of the C-ABI library (include/c2ray_b200.h).

The library is the product: if libc2ray_b200.so is missing or cannot be loaded this module raises
(there is no CPU or PyTorch fallback).  Build it with `python -c "import __graft_entry__ as g; g.build()"`
or `make -C c2ray3dm_b200/csrc`.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("C2B_LIB") or os.path.join(HERE, "libc2ray_b200.so")   # C2B_LIB: development builds (scripts/build_variant.sh)

NUMTAU = 2000
MAX_ITER = 104
UNIQUE_ID_BYTES = 128


class Config(C.Structure):
    _fields_ = [("mesh", C.c_int32 * 3), ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32),
                ("isothermal", C.c_int32), ("type_of_clumping", C.c_int32), ("use_LLS", C.c_int32),
                ("type_of_LLS", C.c_int32), ("subboxsize", C.c_int32), ("max_subbox", C.c_int32),
                ("max_outer_iter", C.c_int32), ("reserved0", C.c_int32)] + \
               [(n, C.c_double) for n in
                "epsilon convergence_fraction minimum_fractional_change minimum_fraction_of_atoms "
                "loss_fraction max_coldensh tau_photo_limit minlogtau dlogtau sigma_HI pi sqrt2 sqrt3 "
                "bh00 albpow colh0 temph0 abu_c "
                "k_B gamma1 minitemp relative_denergy tau_heat_limit H0 Omega0".split()] + \
               [("cosmological", C.c_int32), ("reserved1", C.c_int32)]


class PhotonStats(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                "h0_before h1_before h0_after h1_after totrec totcollisions dh0 total_ion "
                "totalsrc photcons total_photon_loss LLS_loss".split()]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class PassReport(C.Structure):
    _fields_ = [("photon_loss_all", C.c_double), ("sum_nbox_all", C.c_int64), ("updates", C.c_int64),
                ("ms_raytrace", C.c_double), ("ms_allreduce", C.c_double)]


class GlobalReport(C.Structure):
    _fields_ = [("conv_flag", C.c_int32), ("reserved0", C.c_int32), ("min_avg_neutral", C.c_double),
                ("sum_xh_intermed", C.c_double), ("stats", PhotonStats), ("ms_chemistry", C.c_double)]


class StepReport(C.Structure):
    _fields_ = [("niter", C.c_int32), ("converged", C.c_int32), ("conv_criterion", C.c_int32),
                ("reserved0", C.c_int32),
                ("conv_flag", C.c_int32 * MAX_ITER),
                ("rel_change_sum_xh1", C.c_double * MAX_ITER),
                ("rel_change_sum_xh0", C.c_double * MAX_ITER),
                ("photon_loss_all", C.c_double * MAX_ITER),
                ("sum_nbox_all", C.c_int64 * MAX_ITER),
                ("updates", C.c_int64 * MAX_ITER),
                ("iter_stats", PhotonStats * MAX_ITER),
                ("final_stats", PhotonStats),
                ("grtotal_ion", C.c_double), ("grtotal_src", C.c_double),
                ("total_updates", C.c_int64), ("kernel_launches", C.c_int64),
                ("ms_raytrace", C.c_double), ("ms_allreduce", C.c_double),
                ("ms_chemistry", C.c_double), ("ms_total", C.c_double)]


# every symbol include/c2ray_b200.h declares (tests check the .so exports all of them)
SYMBOLS = """c2b_default_config c2b_create c2b_destroy c2b_last_error c2b_device_count
c2b_get_unique_id c2b_comm_init c2b_set_tables c2b_rad_ini_blackbody c2b_set_density
c2b_set_geometry c2b_cosmo_evol c2b_set_clumping_scalar c2b_set_clumping_grid c2b_set_lls_scalar
c2b_set_lls_grid c2b_set_lls_rmax c2b_set_temperature c2b_set_sources c2b_set_xh c2b_evolve3d
c2b_begin_step c2b_pass_all_sources c2b_global_pass c2b_end_step c2b_get_xh c2b_get_xh_av
c2b_get_xh_intermed c2b_get_phih c2b_get_phih_f32 c2b_get_source_nbox c2b_get_source_loss
c2b_get_iter_state c2b_set_iter_state c2b_dev_ptr c2b_synchronize c2b_trace_source_debug
c2b_measure_dfma_rate c2b_save_xh_dev c2b_restore_xh_dev c2b_get_route_counts c2b_deal_sources c2b_get_source_owner
c2b_set_clumping_from_density c2b_get_clumping_grid
c2b_set_heat_tables c2b_get_heat_tables c2b_set_cooling_table c2b_set_redshift c2b_set_temperature_grid
c2b_get_temperature_grid c2b_get_phiheat c2b_get_iter_state_thermal c2b_set_iter_state_thermal""".split()

_lib = None


def load():
    """Loads libc2ray_b200.so; raises RuntimeError when the CUDA extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "c2ray3dm_b200: %s is missing -- the CUDA library is the product and there is no fallback; "
            "build it with __graft_entry__.build() or `make -C c2ray3dm_b200/csrc`" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32)
    L.c2b_default_config.argtypes = [C.POINTER(Config)]
    L.c2b_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.c2b_destroy.argtypes = [vp]
    L.c2b_destroy.restype = None
    L.c2b_last_error.argtypes = [vp]
    L.c2b_last_error.restype = C.c_char_p
    L.c2b_device_count.argtypes = []
    L.c2b_get_unique_id.argtypes = [vp]
    L.c2b_comm_init.argtypes = [vp, vp]
    L.c2b_set_tables.argtypes = [vp, dp, dp, C.c_int32]
    L.c2b_rad_ini_blackbody.argtypes = [vp] + [C.c_double] * 9 + [dp, dp]
    L.c2b_set_density.argtypes = [vp, fp]
    L.c2b_set_geometry.argtypes = [vp, dp, C.c_double]
    L.c2b_cosmo_evol.argtypes = [vp, C.c_double]
    L.c2b_set_clumping_scalar.argtypes = [vp, C.c_float]
    L.c2b_set_clumping_grid.argtypes = [vp, fp]
    L.c2b_set_lls_scalar.argtypes = [vp, C.c_double]
    L.c2b_set_lls_grid.argtypes = [vp, fp]
    L.c2b_set_lls_rmax.argtypes = [vp, C.c_double]
    L.c2b_set_temperature.argtypes = [vp, C.c_double]
    L.c2b_set_sources.argtypes = [vp, C.c_int32, ip, dp, C.c_double]
    L.c2b_set_xh.argtypes = [vp, dp]
    L.c2b_evolve3d.argtypes = [vp, C.c_double, C.c_double, C.c_int32, C.POINTER(StepReport)]
    L.c2b_begin_step.argtypes = [vp, dp]
    L.c2b_pass_all_sources.argtypes = [vp, C.c_int32, C.c_double, C.POINTER(PassReport)]
    L.c2b_global_pass.argtypes = [vp, C.c_double, C.POINTER(GlobalReport)]
    L.c2b_end_step.argtypes = [vp, C.c_double, C.c_int32, C.POINTER(PhotonStats)]
    for f in ("c2b_get_xh", "c2b_get_xh_av", "c2b_get_xh_intermed", "c2b_get_phih"):
        getattr(L, f).argtypes = [vp, dp]
    L.c2b_get_phih_f32.argtypes = [vp, fp]
    L.c2b_get_source_nbox.argtypes = [vp, ip]
    L.c2b_get_source_loss.argtypes = [vp, dp]
    L.c2b_get_iter_state.argtypes = [vp, ip, dp, dp, dp, dp]
    L.c2b_set_iter_state.argtypes = [vp, C.c_int32, C.c_double, dp, dp, dp]
    L.c2b_dev_ptr.argtypes = [vp, C.c_char_p]
    L.c2b_dev_ptr.restype = vp
    L.c2b_synchronize.argtypes = [vp]
    L.c2b_save_xh_dev.argtypes = [vp]
    L.c2b_set_heat_tables.argtypes = [vp, dp, dp, C.c_int32]
    L.c2b_get_heat_tables.argtypes = [vp, dp, dp]
    L.c2b_set_cooling_table.argtypes = [vp, dp, dp, C.c_int32]
    L.c2b_set_redshift.argtypes = [vp, C.c_double]
    L.c2b_set_temperature_grid.argtypes = [vp, fp]
    L.c2b_get_temperature_grid.argtypes = [vp, fp]
    L.c2b_get_phiheat.argtypes = [vp, dp]
    L.c2b_get_iter_state_thermal.argtypes = [vp, dp, fp]
    L.c2b_set_iter_state_thermal.argtypes = [vp, dp, fp]
    L.c2b_restore_xh_dev.argtypes = [vp]
    L.c2b_trace_source_debug.argtypes = [vp, C.c_int32, dp, dp, ip, dp]
    L.c2b_measure_dfma_rate.argtypes = [vp, dp]
    L.c2b_get_route_counts.argtypes = [vp, C.POINTER(C.c_int64)]
    L.c2b_deal_sources.argtypes = [C.c_int32, C.POINTER(C.c_int64), C.c_int32, dp, ip]
    L.c2b_get_source_owner.argtypes = [vp, ip]
    L.c2b_set_clumping_from_density.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_double]
    L.c2b_get_clumping_grid.argtypes = [vp, fp]
    _lib = L
    return L


def default_config():
    cfg = Config()
    rc = load().c2b_default_config(C.byref(cfg))
    if rc:
        raise RuntimeError("c2b_default_config failed")
    return cfg
