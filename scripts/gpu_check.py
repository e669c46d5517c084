"""Developer diagnostic: GPU vs oracle on a few seeded problems, printing error levels."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from problems import make_problem, setup_oracle, setup_gpu
from oracle import oracle as O


def relmax(a, b):
    m = np.abs(b) > 0
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m])))


def main():
    thick, thin, _ = O.rad_ini()
    cases = [
        dict(N=16, nsrc=1, seed=1, state="ionized", use_LLS=False),
        dict(N=21, nsrc=3, seed=5, state="ionized", use_LLS=False),
        dict(N=24, nsrc=3, seed=5, state="ionized", use_LLS=True),
        dict(N=(16, 20, 12), nsrc=3, seed=5, state="ionized", use_LLS=False),
        dict(N=32, nsrc=8, seed=7, state="random", use_LLS=True),
        dict(N=32, nsrc=8, seed=7, state="random", use_LLS=True, type_of_LLS=2, clumping="grid"),
        dict(N=32, nsrc=8, seed=7, state="ionized", use_LLS=True, type_of_LLS=3),
        dict(N=64, nsrc=20, seed=9, state="random", use_LLS=True),
    ]
    for c in cases:
        p = make_problem(**c)
        if c["state"] == "random":
            p["xh"] = 1 - (1 - p["xh"]) * 1e-2
        e = setup_gpu(p)
        gt, gn = e.rad_ini()
        print("tables rel err thick %.2e thin %.2e" % (relmax(gt, thick), relmax(gn, thin)))
        o = setup_oracle(p, tables=(gt, gn))
        # single pass of all sources
        o.xh_av[...] = p["xh"]
        o.set_rates_to_zero()
        t0 = time.time()
        r = o.pass_all_sources()
        t_cpu = time.time() - t0
        e.begin_step()
        g = e.pass_all_sources()
        ph = e.phih_grid
        print(c)
        print("  loss cpu %.15e gpu %.15e  nbox %d %d  updates %d %d  cpu %.3fs gpu %.3fms" % (
            r.photon_loss_all, g.photon_loss_all, r.sum_nbox_all, g.sum_nbox_all, r.updates, g.updates, t_cpu,
            g.ms_raytrace))
        print("  phih max rel err %.3e   (max %.3e)" % (relmax(ph, o.phih), o.phih.max()))
        # single source debug: coldensh_out
        o2 = setup_oracle(p, tables=(gt, gn))
        o2.xh_av[...] = p["xh"]
        o2.set_rates_to_zero()
        rr = o2.do_source(1)
        cd, ph1, nbox, loss = e.trace_source_debug(1)
        print("  src1: nbox %d %d loss %.15e %.15e coldens relerr %.3e zero-mismatch %d phih relerr %.3e" % (
            rr.nbox, nbox, rr.photon_loss_src, loss, relmax(cd, o2.coldensh_out),
            int(np.sum((cd == 0) != (o2.coldensh_out == 0))), relmax(ph1, o2.phih)))
        # full step
        o3 = setup_oracle(p, tables=(gt, gn))
        dt = 1e7 * 3.15576e7 / 10
        t0 = time.time()
        ro = o3.evolve3D(dt)
        t_cpu = time.time() - t0
        e.set_xh(p["xh"])
        rg = e.evolve3D(0.0, dt)
        print("  evolve3D: niter %d %d conv %d %d  updates %d %d  cpu %.2fs gpu %.2fms (rt %.2f chem %.2f) launches %d" % (
            ro.niter, rg.niter, ro.converged, rg.converged, ro.total_updates, rg.total_updates, t_cpu, rg.ms_total,
            rg.ms_raytrace, rg.ms_chemistry, rg.kernel_launches))
        print("  xh abs err %.3e  xh_av abs err %.3e  phih rel %.3e" % (
            np.max(np.abs(e.xh - o3.xh)), np.max(np.abs(e.xh_av - o3.xh_av)), relmax(e.phih_grid, o3.phih)))
        print("  conv_flag", list(ro.conv_flag[1:ro.niter + 1]), list(rg.conv_flag[1:rg.niter + 1]))
        so, sg = ro.final_stats, rg.final_stats
        for n in ("totrec", "totcollisions", "dh0", "total_ion", "photcons", "total_photon_loss"):
            a, b = getattr(so, n), getattr(sg, n)
            print("    %-18s %.12e %.12e rel %.2e" % (n, a, b, abs(a - b) / max(abs(a), 1e-300)))
        e.close()
    e = setup_gpu(make_problem(16))
    print("DFMA rate %.3e instr/s" % e.measure_dfma_rate())


if __name__ == "__main__":
    main()
