"""Timing study: repeated pass_all_sources on a fixed mid-reionization state (no chemistry in between), so
kernel changes can be compared on identical work:  python scripts/time_pass.py [mesh] [nsrc]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from c2ray3dm_b200 import Evolve

mesh = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
w = bench.build_workload(mesh, nsrc, 25.0)
e = Evolve(mesh, type_of_clumping=5, use_LLS=True)
e.rad_ini(); e.set_geometry(w["dr"], w["vol"]); e.set_clumping(w["clumping"]); e.set_LLS(coldensh_LLS=w["coldensh_LLS"])
e.set_sources(w["srcpos"], w["normflux"]); e.set_density(w["ndens"]); e.set_xh(w["xh"])
e.evolve3D(0.0, 0.5e6 * 3.15576e7)
e.begin_step()
for rep in range(3):
    r = e.pass_all_sources()
    print("pass %d: %.1f ms, %.2f G updates/s (%d updates)" % (rep, r.ms_raytrace, r.updates / r.ms_raytrace / 1e6, r.updates), flush=True)
