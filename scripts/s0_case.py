"""Early-reionization (S0) case of the bench workload for profiling: xh = 2e-4 everywhere, one evolve3D call after a
warm-up call.  usage: python scripts/s0_case.py [mesh] [nsrc]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
from c2ray3dm_b200 import Evolve, constants as K

mesh = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
w = bench._build_workload(mesh, nsrc, 25.0)
e = Evolve(mesh, type_of_clumping=5, use_LLS=True, type_of_LLS=1)
e.rad_ini()
e.set_geometry(w["dr"], w["vol"])
e.set_clumping(w["clumping"])
e.set_LLS(coldensh_LLS=w["coldensh_LLS"])
e.set_sources(w["srcpos"], w["normflux"])
e.set_density(w["ndens"])
xh0 = np.full(mesh ** 3, K.xh_initial)
dt = 0.5e6 * 3.15576e7
for i in range(2):
    e.set_xh(xh0)
    e.synchronize()
    t0 = time.perf_counter()
    rep = e.evolve3D(0.0, dt)
    e.synchronize()
    t = time.perf_counter() - t0
    print("S0 step %d: niter %d updates %d  %.2f ms  raytrace %.2f ms  chemistry %.2f ms  routes %s" % (
        i, rep.niter, rep.total_updates, 1e3 * t, rep.ms_raytrace, rep.ms_chemistry, e.route_counts()))
e.close()
