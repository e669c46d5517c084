"""The compiled-language host above the C ABI (host/evolve.cpp: `evolve::evolve3D(time,dt,restart)` with the
reference's module state, the C++ twin of fortran/evolve_b200.F90) driven by host/run_case.cpp, which plays
C2Ray.F90:352-394: cosmo_evol on the host copies, call evolve3D, keep xh for the next step."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from problems import make_problem, setup_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "host", "run_case")
DT = 1e6 * 3.15576e7


def _write_case(path, p, tables, nsteps, zfactor):
    m = p["mesh"]
    with open(path, "wb") as f:
        f.write(struct.pack("<8i", m[0], m[1], m[2], nsteps, len(p["normflux"]), p["type_of_clumping"],
                            int(p["use_LLS"]), p["type_of_LLS"]))
        f.write(struct.pack("<13d", DT, p["dr"][0], p["dr"][1], p["dr"][2], p["vol"], p["temper"], p["clumping"],
                            p["coldensh_LLS"], p["R_max_LLS"], p["S_star"], zfactor, 0.0, 0.0))
        f.write(np.ascontiguousarray(p["ndens"], dtype=np.float32).tobytes())
        f.write(np.ascontiguousarray(p["xh"], dtype=np.float64).tobytes())
        if p["type_of_clumping"] >= 3:
            f.write(np.ascontiguousarray(p["clumping_grid"], dtype=np.float32).tobytes())
        if p["use_LLS"] and p["type_of_LLS"] == 2:
            f.write(np.ascontiguousarray(p["LLS_grid"], dtype=np.float32).tobytes())
        f.write(np.ascontiguousarray(p["srcpos"], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(p["normflux"], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(tables[0], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(tables[1], dtype=np.float64).tobytes())


def test_driver_fails_loudly_without_gpu(tmp_path):
    """no CUDA device => evolve3D reports the library's error and the driver exits non-zero (no fallback)"""
    from c2ray3dm_b200 import lib
    if lib.load().c2b_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from oracle import oracle as O
    p = make_problem(8, nsrc=1, seed=1)
    case = tmp_path / "case.bin"
    _write_case(str(case), p, O.rad_ini()[:2], 1, 1.0)
    r = subprocess.run([DRV, str(case), str(tmp_path / "out.bin"), str(tmp_path / "log.txt")], capture_output=True, text=True)
    assert r.returncode == 1
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("case", [
    dict(N=24, nsrc=5, seed=41, state="random", use_LLS=True),
    dict(N=(20, 16, 24), nsrc=4, seed=42, state="random", use_LLS=True, type_of_LLS=2, clumping="grid"),
], ids=["cubic_lls1", "noncubic_lls2_clumpgrid"])
def test_cpp_host_history_matches_oracle(case, tmp_path):
    from oracle import oracle as O
    tables = O.rad_ini()[:2]
    p = make_problem(**case)
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    nsteps, zf = 3, 1.004
    cfile, ofile, lfile = tmp_path / "case.bin", tmp_path / "out.bin", tmp_path / "log.txt"
    _write_case(str(cfile), p, tables, nsteps, zf)
    r = subprocess.run([DRV, str(cfile), str(ofile), str(lfile)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    n = p["mesh"][0] * p["mesh"][1] * p["mesh"][2]
    raw = open(ofile, "rb").read()
    rec = 4 + 6 * 8 + 2 * 8 * n
    assert len(raw) == nsteps * rec
    # the oracle through the same host sequence
    o = setup_oracle(p, tables=tables)
    ndens = p["ndens"].copy()
    dr, vol = p["dr"].copy(), p["vol"]
    for step in range(nsteps):
        z3 = zf * zf * zf
        dr, vol = dr * zf, vol * z3
        ndens = (ndens.astype(np.float64) / z3).astype(np.float32)
        o.set_density(ndens)
        o.set_geometry(dr, vol)
        ro = o.evolve3D(DT)
        off = step * rec
        niter = struct.unpack_from("<i", raw, off)[0]
        st = struct.unpack_from("<6d", raw, off + 4)
        xh = np.frombuffer(raw, dtype=np.float64, count=n, offset=off + 52).reshape(p["shape"])
        ph = np.frombuffer(raw, dtype=np.float64, count=n, offset=off + 52 + 8 * n).reshape(p["shape"])
        assert niter == ro.niter
        np.testing.assert_allclose(xh, o.xh, rtol=0, atol=1e-6)
        nz = o.phih != 0
        assert np.max(np.abs(ph[nz] - o.phih[nz]) / o.phih[nz]) <= 1e-6
        assert st[0] == pytest.approx(ro.final_stats.total_ion, rel=1e-6, abs=1e-6 * abs(ro.final_stats.totrec))
        assert st[1] == pytest.approx(ro.final_stats.totrec, rel=1e-6)
        assert st[2] == pytest.approx(ro.final_stats.totcollisions, rel=1e-6)
        assert st[4] == pytest.approx(ro.grtotal_ion, rel=1e-6, abs=1e-6 * abs(ro.final_stats.totrec))
        assert st[5] == pytest.approx(ro.grtotal_src, rel=1e-12)
    log = open(lfile).read()
    for line in ("Convergence tests:", "Doing all sources", "Average number of subboxes:", "Doing global",
                 "Number of non-converged points:", "Multiple sources convergence reached"):
        assert line in log      # the log lines of evolve.F90 survive the swap


def _read_fortran_records(path, dtypes_counts):
    """reads the records of a Fortran sequential unformatted file (4-byte length markers)"""
    raw = open(path, "rb").read()
    off, out = 0, []
    for dt, cnt in dtypes_counts:
        n0 = struct.unpack_from("<I", raw, off)[0]
        assert n0 == np.dtype(dt).itemsize * cnt
        out.append(np.frombuffer(raw, dtype=dt, count=cnt, offset=off + 4))
        assert struct.unpack_from("<I", raw, off + 4 + n0)[0] == n0
        off += 8 + n0
    assert off == len(raw)
    return out


@pytest.mark.gpu
def test_cpp_host_dump_and_restart_match_oracle(tmp_path):
    """write_iteration_dump / start_from_dump of the host (evolve.F90:285-426): a run that dumps after every
    pass_all_sources leaves iterdump1.bin / iterdump2.bin with the reference's record layout (niter |
    photon_loss_all | phih_grid | xh_av | xh_intermed); the records match the oracle's dump of the same iteration,
    and a second process restarted from the file (restart = 1 or 2) ends exactly where the uninterrupted run did"""
    from oracle import oracle as O
    tables = O.rad_ini()[:2]
    p = make_problem(N=24, nsrc=5, seed=41, state="random", use_LLS=True)
    p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    n = 24 ** 3
    cfile, lfile = tmp_path / "case.bin", tmp_path / "log.txt"
    _write_case(str(cfile), p, tables, 1, 1.0)
    # uninterrupted run, dumping after every pass
    r = subprocess.run([DRV, str(cfile), str(tmp_path / "full.bin"), str(lfile), str(tmp_path), "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    raw = open(tmp_path / "full.bin", "rb").read()
    niter_full = struct.unpack_from("<i", raw, 0)[0]
    xh_full = np.frombuffer(raw, dtype=np.float64, count=n, offset=52)
    assert niter_full >= 3
    layout = [(np.int32, 1), (np.float64, 1), (np.float64, n), (np.float64, n), (np.float64, n)]
    d1 = _read_fortran_records(tmp_path / "iterdump1.bin", layout)
    d2 = _read_fortran_records(tmp_path / "iterdump2.bin", layout)
    # the two files hold the last odd-numbered and the last even-numbered dump
    its = sorted([int(d1[0][0]), int(d2[0][0])])
    assert its == [niter_full - 1, niter_full]
    # the oracle's dump of the same iteration
    for d in (d1, d2):
        k = int(d[0][0])
        o = setup_oracle(p, tables=tables)
        o.set_dump_iteration(k)
        ro = o.evolve3D(DT)
        assert ro.niter == niter_full
        nit, pl, ph, xav, xint = o.get_dump()
        assert nit == k and float(d[1][0]) == pytest.approx(pl, rel=1e-6)
        nz = ph.reshape(-1) != 0
        assert np.max(np.abs(d[2][nz] - ph.reshape(-1)[nz]) / ph.reshape(-1)[nz]) <= 1e-6
        np.testing.assert_allclose(d[3], xav.reshape(-1), rtol=0, atol=1e-6)
        np.testing.assert_allclose(d[4], xint.reshape(-1), rtol=0, atol=1e-6)
    # restart a fresh process from the older of the two dumps (the other one is the final iteration)
    which = 1 if int(d1[0][0]) == niter_full - 1 else 2
    r = subprocess.run([DRV, str(cfile), str(tmp_path / "rest.bin"), str(tmp_path / "log2.txt"), str(tmp_path),
                        "1e9", str(which)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    raw2 = open(tmp_path / "rest.bin", "rb").read()
    assert struct.unpack_from("<i", raw2, 0)[0] == niter_full
    xh_rest = np.frombuffer(raw2, dtype=np.float64, count=n, offset=52)
    np.testing.assert_allclose(xh_rest, xh_full, rtol=0, atol=1e-12)
    log2 = open(tmp_path / "log2.txt").read()
    assert "Read iteration %d from dump file" % (niter_full - 1) in log2
    assert "Multiple sources convergence reached" in log2
