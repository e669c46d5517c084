"""Generates tests/golden/*.npz from the CPU oracle (oracle/c2ray_oracle.c).

The reference ships no golden vectors and cannot be built here (no Fortran compiler), so these
fixtures pin the ORACLE's behaviour (drift guard) and give the GPU tests a committed target; they
are not reference outputs.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from problems import make_problem, setup_oracle
from oracle import oracle as O

CASES = {
    "g16_lls": dict(N=16, nsrc=4, seed=11, state="random", use_LLS=True),
    "g20_clump": dict(N=20, nsrc=5, seed=12, state="random", use_LLS=True, clumping="grid"),
    "g12x16x10": dict(N=(12, 16, 10), nsrc=3, seed=13, state="ionized", use_LLS=False),
}
DT = 1e6 * 3.15576e7


def make_case(name):
    c = CASES[name]
    p = make_problem(**c)
    if c["state"] == "random":
        p["xh"] = 1 - (1 - p["xh"]) * 1e-2
    return p


def main():
    thick, thin, diag = O.rad_ini()
    idx = np.array([0, 1, 500, 1000, 1500, 1668, 1700, 1800, 1900, 2000])
    np.savez(os.path.join(HERE, "tables.npz"), idx=idx, thick=thick[idx], thin=thin[idx],
             thick_sum=thick.sum(), thin_sum=thin.sum(), romw7=np.array(diag.romw7),
             S_star_unscaled=diag.S_star_unscaled)
    for name in CASES:
        p = make_case(name)
        o = setup_oracle(p)
        o.xh_av[...] = p["xh"]
        o.set_rates_to_zero()
        r = o.pass_all_sources()
        phih_pass = o.phih.copy()
        o2 = setup_oracle(p)
        rep = o2.evolve3D(DT)
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            phih_pass=phih_pass, photon_loss_all=r.photon_loss_all, sum_nbox=r.sum_nbox_all,
                            updates=r.updates, xh=o2.xh.copy(), xh_av=o2.xh_av.copy(), phih=o2.phih.copy(),
                            niter=rep.niter, conv_flag=np.array(rep.conv_flag[:rep.niter + 1]),
                            photcons=rep.final_stats.photcons, total_ion=rep.final_stats.total_ion,
                            totrec=rep.final_stats.totrec, totcollisions=rep.final_stats.totcollisions)
        print(name, "niter", rep.niter, "updates", r.updates)


if __name__ == "__main__":
    main()
