import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from c2ray3dm_b200 import Evolve, synthetic as syn

mesh = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
for bubble, cf in ((25.0, 4.0), (25.0, 0.5), (15.0, 0.5)):
    w = bench.build_workload(mesh, nsrc, bubble)
    clump = (1.0 + cf * (w["ndens"].astype(np.float64) / syn.avg_dens(9.0))).astype(np.float32)
    e = Evolve(mesh, type_of_clumping=5, use_LLS=True)
    e.rad_ini(); e.set_geometry(w["dr"], w["vol"]); e.set_clumping(clump); e.set_LLS(coldensh_LLS=w["coldensh_LLS"])
    e.set_sources(w["srcpos"], w["normflux"]); e.set_density(w["ndens"]); e.set_xh(w["xh"])
    dt = 0.5e6 * 3.15576e7
    for step in range(3):
        t = time.time(); rep = e.evolve3D(0, dt); t = time.time() - t
        print("bubble", bubble, "cf", cf, "step", step, "niter", rep.niter, "updates/iter", [int(u) for u in rep.updates[1:rep.niter+1]],
              "ms rt %.1f chem %.1f total %.1f wall %.2fs" % (rep.ms_raytrace, rep.ms_chemistry, rep.ms_total, t),
              "G/s %.2f" % (rep.total_updates / rep.ms_raytrace / 1e6), "mean x %.3f" % (e.xh.mean()), "photcons %.3f" % rep.final_stats.photcons, flush=True)
    e.close()
