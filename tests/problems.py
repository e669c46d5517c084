"""Seeded small problems shared by the oracle tests and the GPU parity tests."""
import numpy as np

from c2ray3dm_b200 import constants as K
from c2ray3dm_b200 import synthetic as syn


def make_problem(N, nsrc=3, seed=1, zred=9.0, state="random", use_LLS=True, type_of_LLS=1,
                 clumping="scalar", flux=1e7, srcpos=None, dens="random", boxsize=None):
    """Returns a dict with everything both implementations need.  N may be an int or (n1,n2,n3)."""
    mesh = (N, N, N) if np.isscalar(N) else tuple(N)
    shape = (mesh[2], mesh[1], mesh[0])
    rng = np.random.default_rng(seed)
    if boxsize is None:
        boxsize = K.boxsize_test
    dr1 = syn.comoving_dr(mesh[0], boxsize) / (1.0 + zred)
    dr = np.array([dr1, dr1, dr1])
    vol = dr1 * dr1 * dr1
    nbar = syn.avg_dens(zred)
    if dens == "uniform":
        ndens = np.full(shape, nbar, dtype=np.float32)
    else:
        ndens = (nbar * np.exp(rng.normal(0.0, 0.7, size=shape))).astype(np.float32)
    if state == "random":
        xh = syn.random_state(shape, seed + 100)
    elif state == "neutral":
        xh = np.full(shape, K.xh_initial)
    elif state == "ionized":
        xh = np.full(shape, 1.0 - 1e-5)
    else:
        raise ValueError(state)
    if srcpos is None:
        srcpos = np.stack([rng.integers(1, mesh[d] + 1, size=nsrc) for d in range(3)], axis=1).astype(np.int32)
    else:
        srcpos = np.asarray(srcpos, dtype=np.int32).reshape(-1, 3)
        nsrc = srcpos.shape[0]
    normflux = flux * 10.0 ** rng.uniform(-1.5, 0.0, size=nsrc)
    p = dict(mesh=mesh, shape=shape, dr=dr, vol=vol, ndens=ndens, xh=xh, srcpos=srcpos, normflux=normflux,
             S_star=K.bb_S_star, use_LLS=use_LLS, type_of_LLS=type_of_LLS, temper=K.initial_temperature,
             zred=zred)
    p["coldensh_LLS"] = syn.lls_coldens(dr1, zred)
    p["LLS_grid"] = None
    p["R_max_LLS"] = 0.0
    if use_LLS and type_of_LLS == 2:
        p["LLS_grid"] = (p["coldensh_LLS"] * rng.uniform(0.5, 2.0, size=shape)).astype(np.float32)
    if use_LLS and type_of_LLS == 3:
        p["R_max_LLS"] = 0.3 * mesh[0] * dr1
    if clumping == "scalar":
        p["type_of_clumping"], p["clumping"], p["clumping_grid"] = 1, 1.0, None
    elif clumping == "scalar2":
        p["type_of_clumping"], p["clumping"], p["clumping_grid"] = 2, 7.25, None
    else:
        p["type_of_clumping"], p["clumping"] = 5, 1.0
        p["clumping_grid"] = syn.clumping_from_density(ndens, zred)
    return p


def setup_oracle(p, tables=None):
    from oracle import oracle as O
    o = O.Oracle(p["mesh"])
    if tables is not None:
        o.set_tables(*tables)
    o.set_density(p["ndens"])
    o.set_geometry(p["dr"], p["vol"])
    o.set_clumping(p["type_of_clumping"], p["clumping"], p["clumping_grid"])
    o.set_lls(p["use_LLS"], p["type_of_LLS"], p["coldensh_LLS"], p["LLS_grid"], p["R_max_LLS"])
    o.set_temperature(p["temper"])
    o.set_sources(p["srcpos"], p["normflux"], p["S_star"])
    o.set_xh(p["xh"])
    return o


def setup_gpu(p, rank=0, nranks=1, device=0, tables=None, **overrides):
    from c2ray3dm_b200 import Evolve
    e = Evolve(p["mesh"], device=device, rank=rank, nranks=nranks, type_of_clumping=p["type_of_clumping"],
               use_LLS=p["use_LLS"], type_of_LLS=p["type_of_LLS"], **overrides)
    if tables is None:
        e.rad_ini()
    else:
        e.set_tables(*tables)
    e.set_density(p["ndens"])
    e.set_geometry(p["dr"], p["vol"])
    if p["clumping_grid"] is not None:
        e.set_clumping(p["clumping_grid"])
    else:
        e.set_clumping(p["clumping"])
    e.set_LLS(coldensh_LLS=p["coldensh_LLS"], LLS_grid=p["LLS_grid"], R_max_LLS=p["R_max_LLS"])
    e.set_temperature(p["temper"])
    e.set_sources(p["srcpos"], p["normflux"], p["S_star"])
    e.set_xh(p["xh"])
    return e


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor + 1e-300))
