"""Small evolve3D runs for compute-sanitizer (memcheck / racecheck): the three ray-trace work-group shapes (one CTA,
one cluster, one warp per source with its hand-over to the one-CTA kernel), LLS modes 1 and 2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from problems import make_problem, setup_gpu
for mode in ("cta", "cluster"):
    os.environ["C2B_CLUSTER_MIN_NBOX"] = "100000" if mode == "cta" else "0"
    os.environ["C2B_CLUSTER_MAX_SOURCES"] = "1000000"
    for case in (dict(N=24, nsrc=3, seed=5, state="ionized", use_LLS=True),
                 dict(N=(16, 20, 12), nsrc=2, seed=5, state="ionized", use_LLS=True, type_of_LLS=2, clumping="grid")):
        p = make_problem(**case)
        e = setup_gpu(p)
        rep = e.evolve3D(0.0, 3.15576e13)
        print(mode, case["N"], "niter", rep.niter, "updates", rep.total_updates, flush=True)
        e.close()

# one warp per source (early reionization), with sources that outgrow the first subbox and are handed over
os.environ["C2B_CLUSTER_MIN_NBOX"] = "100000"
os.environ["C2B_WARP_MIN_SOURCES"] = "1"
import numpy as np
for lls, clump in ((1, "scalar"), (2, "grid")):
    p = make_problem(24, nsrc=30, seed=31, state="neutral", use_LLS=True, type_of_LLS=lls, clumping=clump, flux=2e6)
    p["normflux"][::7] *= 3000.0
    e = setup_gpu(p)
    for step in range(2):
        rep = e.evolve3D(0.0, 3.15576e13)
    print("warp", lls, "niter", rep.niter, "updates", rep.total_updates, "routes", e.route_counts(), flush=True)
    e.close()
