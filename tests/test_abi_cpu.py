"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol the
header declares, its default configuration carries the reference's literal semantics, and it fails
loudly (no fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import c2ray3dm_b200 as pkg
from c2ray3dm_b200 import lib as L
from c2ray3dm_b200 import constants as K
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "c2ray_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(c2b_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    syms = _header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), "libc2ray_b200.so does not export %s" % s
    assert sorted(L.SYMBOLS) == syms


def test_default_config_matches_reference_literals():
    cfg = L.default_config()
    c = O.constants()
    assert list(cfg.mesh) == [300, 300, 300]          # sizes.f90:33
    assert (cfg.isothermal, cfg.type_of_clumping, cfg.use_LLS, cfg.type_of_LLS) == (1, 1, 1, 1)
    assert (cfg.subboxsize, cfg.max_subbox, cfg.max_outer_iter) == (5, 1000, 100)
    for name_cfg, name_c in [("epsilon", "epsilon"), ("convergence_fraction", "convergence_fraction"),
                             ("minimum_fractional_change", "minimum_fractional_change"),
                             ("minimum_fraction_of_atoms", "minimum_fraction_of_atoms"),
                             ("loss_fraction", "loss_fraction"), ("max_coldensh", "max_coldensh"),
                             ("tau_photo_limit", "tau_photo_limit"), ("minlogtau", "minlogtau"),
                             ("dlogtau", "dlogtau"), ("sigma_HI", "sigma_HI_at_ion_freq"), ("pi", "pi"),
                             ("sqrt2", "sqrt2"), ("sqrt3", "sqrt3"), ("bh00", "bh00"), ("albpow", "albpow"),
                             ("colh0", "colh0"), ("temph0", "temph0"), ("abu_c", "abu_c")]:
        assert getattr(cfg, name_cfg) == getattr(c, name_c), name_cfg


def test_python_constants_match_oracle():
    c = O.constants()
    for n in ("pi", "sigma_HI_at_ion_freq", "eth0", "ev2k", "temph0", "colh0", "ev2fr", "ion_freq_HI",
              "ion_freq_HeII", "two_pi_over_c_square", "bh00", "albpow", "hplanck", "k_B", "m_p", "abu_he",
              "abu_c", "mu", "h", "Omega0", "Omega_B", "Mpc", "H0", "rho_crit_0", "YEAR", "R_SOLAR",
              "xh_initial", "bb_Teff", "bb_S_star", "bb_MaxFreq"):
        assert getattr(K, n) == getattr(c, n), n


def test_struct_layouts_match_header_sizes():
    # sizes implied by the header (no padding surprises between ctypes and the C structs)
    assert C.sizeof(L.Config) == 14 * 4 + 25 * 8 + 2 * 4
    cfg = L.default_config()
    assert cfg.cosmological == 1 and cfg.k_B == 1.381e-16 and abs(cfg.gamma1 - 2.0 / 3.0) < 1e-15 and cfg.minitemp == 1.0
    assert C.sizeof(L.PhotonStats) == 12 * 8
    assert C.sizeof(L.PassReport) == 5 * 8
    assert C.sizeof(L.GlobalReport) == 8 + 2 * 8 + 12 * 8 + 8
    n = L.MAX_ITER
    assert C.sizeof(L.StepReport) == 16 + n * 4 + 5 * n * 8 + n * 96 + 96 + 2 * 8 + 2 * 8 + 4 * 8


@pytest.mark.skipif(L.load().c2b_device_count() > 0, reason="a CUDA device is present")
def test_create_fails_loudly_without_gpu():
    with pytest.raises(pkg.C2RayError) as ei:
        pkg.Evolve(16)
    assert "no CPU fallback" in str(ei.value)


def test_bad_config_is_rejected_before_touching_cuda():
    lib = L.load()
    cfg = L.default_config()
    h = C.c_void_p()
    cfg.mesh[0] = 2
    assert lib.c2b_create(C.byref(cfg), C.byref(h)) == 101 and not h.value
    cfg = L.default_config()
    cfg.mesh[0], cfg.mesh[1], cfg.mesh[2] = 2048, 2048, 2048      # 2^33 cells: default-integer cell counts would overflow
    assert lib.c2b_create(C.byref(cfg), C.byref(h)) == 101 and b"2^31" in lib.c2b_last_error(None)
    cfg = L.default_config()
    cfg.rank, cfg.nranks = 3, 2
    assert lib.c2b_create(C.byref(cfg), C.byref(h)) == 103
    assert lib.c2b_create(None, C.byref(h)) == 100
    # null handles are refused, not dereferenced
    assert lib.c2b_set_density(None, None) == 100
    assert lib.c2b_evolve3d(None, 0.0, 1.0, 0, None) == 100
    lib.c2b_destroy(None)


def test_shard_sources_is_round_robin():
    """do ns1=1+rank,NumSrc,npr (master_slave.F90:85)"""
    assert pkg.shard_sources(10, 0, 4) == [1, 5, 9]
    assert pkg.shard_sources(10, 3, 4) == [4, 8]
    assert pkg.shard_sources(1, 1, 2) == []
    allsrc = sorted(sum((pkg.shard_sources(23, r, 8) for r in range(8)), []))
    assert allsrc == list(range(1, 24))


def test_synthetic_generators_are_deterministic():
    from c2ray3dm_b200 import synthetic as syn
    a = syn.lognormal_density(16, 9.0, 20240607)
    b = syn.lognormal_density(16, 9.0, 20240607)
    assert a.dtype == np.float32 and np.array_equal(a, b)
    assert a.mean() == pytest.approx(syn.avg_dens(9.0), rel=0.2)
    pos, nf = syn.sources_at_density_peaks(a, 10)
    assert pos.shape == (10, 3) and pos.min() >= 1 and pos.max() <= 16
    assert nf[0] == pytest.approx(1e7) and np.all(np.diff(nf) <= 0)
    i, j, k = pos[0]
    assert a[k - 1, j - 1, i - 1] == a.max()
    assert syn.avg_dens(9.0) == pytest.approx(1.9811847154954507e-4, rel=1e-12)
    assert syn.comoving_dr(128) == pytest.approx(3.4441966103425106e24, rel=1e-12)


def test_deal_sources_rule():
    """the host rule that deals sources to ranks between passes (multi-GPU load balance): every source to exactly
    one rank, shares of the predicted cost proportional to the ranks' speeds, deterministic"""
    lib = L.load()
    rng = np.random.default_rng(3)
    n = 5000
    cost = (rng.integers(1, 40, size=n).astype(np.int64)) ** 3
    i64 = C.POINTER(C.c_int64)
    i32 = C.POINTER(C.c_int32)
    dp = C.POINTER(C.c_double)

    def deal(nranks, speed):
        own = np.full(n, -1, dtype=np.int32)
        sp = None if speed is None else np.asarray(speed, dtype=np.float64)
        rc = lib.c2b_deal_sources(n, cost.ctypes.data_as(i64), nranks, None if sp is None else sp.ctypes.data_as(dp),
                                  own.ctypes.data_as(i32))
        assert rc == 0
        return own

    own = deal(8, None)
    assert own.min() == 0 and own.max() == 7
    share = np.array([cost[own == r].sum() for r in range(8)], dtype=np.float64)
    assert np.max(np.abs(share / share.mean() - 1)) < 1e-3          # equal speeds: equal cost within one long trace
    assert np.array_equal(own, deal(8, None))                        # deterministic
    own2 = deal(2, [2.0, 1.0])
    s0, s1 = cost[own2 == 0].sum(), cost[own2 == 1].sum()
    assert s0 / s1 == pytest.approx(2.0, rel=1e-3)                   # twice the speed, twice the work
    # one rank: everything to rank 0; bad arguments are refused
    assert np.all(deal(1, None) == 0)
    assert lib.c2b_deal_sources(n, None, 2, None, own.ctypes.data_as(i32)) != 0
    assert lib.c2b_deal_sources(0, None, 2, None, None) == 0
