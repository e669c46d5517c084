"""BASELINE.json configs on the GPU: config 1/2 (nbody_test sources, uniform density) and the big meshes."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from c2ray3dm_b200 import Evolve, synthetic as syn, constants as K

YEAR = 3.15576e7


def test_problem(N, sources, nsteps, dt_myr=1.0, zred=9.0, label=""):
    """nbody_test.F90 flavour: uniform mean density, 100/h Mpc box, LLS type 1 model 5, xh=2e-4, T=1e4"""
    e = Evolve(N, type_of_clumping=1, use_LLS=True, type_of_LLS=1)
    e.rad_ini()
    dr, vol = syn.proper_geometry(N, zred)
    e.set_geometry(dr, vol)
    e.set_density(syn.uniform_density(N, zred))
    e.set_clumping(1.0)
    e.set_LLS(coldensh_LLS=syn.lls_coldens(dr[0], zred))
    pos = np.array([s[:3] for s in sources], dtype=np.int32)
    nf = np.array([s[3] / 1e48 for s in sources])
    e.set_sources(pos, nf)
    e.set_xh(np.full(N ** 3, K.xh_initial))
    out = []
    for step in range(nsteps):
        t = time.time()
        rep = e.evolve3D(step * dt_myr * 1e6 * YEAR, dt_myr * 1e6 * YEAR)
        wall = time.time() - t
        out.append(dict(step=step, niter=rep.niter, updates=int(rep.total_updates), s=wall, ms_rt=rep.ms_raytrace,
                        photcons=rep.final_stats.photcons, mean_x=float(e.xh.mean())))
        print(label, out[-1], flush=True)
    e.close()
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    std = [(50, 50, 50, 1e55), (51, 50, 50, 1e55), (52, 50, 50, 1e55), (53, 50, 50, 1e55), (20, 10, 10, 1e57),
           (70, 70, 50, 1e55), (72, 70, 50, 1e55), (70, 72, 50, 1e55), (72, 72, 50, 1e56), (20, 10, 90, 1e54)]
    if which in ("all", "c1"):
        test_problem(300, [(50, 50, 50, 1e57)], 10, label="config1 300^3 1 src")
    if which in ("all", "c2"):
        test_problem(128, std, 10, label="config2 128^3 10 src")
