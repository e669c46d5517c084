"""Synthetic inputs for the configurations named in BASELINE.json / SURVEY 8(d).

Everything here is host-side input preparation (what nbody_test.F90, density_module.F90,
sourceprops.F90, LLS.F90 and cosmology.F90 do for the reference before evolve3D is called).
"""
import numpy as np

from . import constants as K


def avg_dens(zred):
    """set_constant_average_density, density_module.F90:129-147"""
    return K.rho_crit_0 * K.Omega_B / (K.mu * K.m_p) * (1.0 + zred) ** 3


def comoving_dr(N, boxsize=K.boxsize_test):
    """grid_ini, grid.F90:108-116: cell size of a `boxsize` Mpc/h box (comoving, cm)"""
    return boxsize * K.Mpc / K.h / float(np.float32(N))


def proper_geometry(N, zred, boxsize=K.boxsize_test):
    """dr(3), vol after cosmo_evol from comoving (zred=0) to zred (cosmology.F90:161-193)"""
    zfactor = 1.0 / (1.0 + zred)
    dr = comoving_dr(N, boxsize) * zfactor
    return np.array([dr, dr, dr]), dr * dr * dr


def lls_coldens(dr1, zred):
    """set_LLS type 1, LLS_model 5 "constant comoving mfp" (LLS.F90:96-99,167-182)"""
    A_LLS, z_ref, yz_LLS = float(np.float32(10.0)), float(np.float32(0.0)), float(np.float32(-1.0))
    mfp = A_LLS * ((1.0 + zred) / (1.0 + z_ref)) ** yz_LLS
    mfp = max(mfp, float(np.float32(1.0)) / (1.0 + zred))
    n_LLS = dr1 / (mfp * K.Mpc)
    N_1 = float(np.float32(1.0)) / K.sigma_HI_at_ion_freq
    return N_1 * n_LLS


def uniform_density(N, zred):
    if np.isscalar(N):
        N = (N, N, N)
    return np.full((N[2], N[1], N[0]), avg_dens(zred), dtype=np.float32)


def lognormal_density(N, zred, seed, sigma=1.0, smooth_cells=2.0):
    """nbar*exp(sigma*g - sigma^2/2), g = unit-variance Gaussian field
    smoothed with a `smooth_cells` Gaussian; float32, i fastest."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((N, N, N), dtype=np.float32)
    k = np.fft.fftfreq(N).astype(np.float32) * 2.0 * np.pi
    kz = k[:, None, None]
    ky = k[None, :, None]
    kx = np.fft.rfftfreq(N).astype(np.float32)[None, None, :] * 2.0 * np.pi
    filt = np.exp(-0.5 * smooth_cells ** 2 * (kx * kx + ky * ky + kz * kz))
    g = np.fft.irfftn(np.fft.rfftn(g) * filt, s=(N, N, N), axes=(0, 1, 2)).astype(np.float64)
    g = (g - g.mean()) / g.std()
    return (avg_dens(zred) * np.exp(sigma * g - 0.5 * sigma * sigma)).astype(np.float32)


def sources_at_density_peaks(ndens, nsrc, max_normflux=1e7):
    """the `nsrc` densest cells (ties by linear index), descending density;
    NormFlux_s = max_normflux * n_s / max(n).  Returns (srcpos 1-based (nsrc,3), normflux)."""
    flat = ndens.reshape(-1)
    if nsrc < flat.size:
        # candidates: every cell at least as dense as the nsrc-th densest (a partition instead of a full sort)
        kth = np.partition(flat, flat.size - nsrc)[flat.size - nsrc]
        cand = np.flatnonzero(flat >= kth)
    else:
        cand = np.arange(flat.size)
    order = cand[np.lexsort((cand, -flat[cand].astype(np.float64)))][:nsrc]
    n3, n2, n1 = ndens.shape
    k, rem = np.divmod(order, n2 * n1)
    j, i = np.divmod(rem, n1)
    srcpos = np.stack([i + 1, j + 1, k + 1], axis=1).astype(np.int32)
    nf = max_normflux * flat[order].astype(np.float64) / float(flat.max())
    return srcpos, nf


def clumping_from_density(ndens, zred, slope=0.5):
    """a deterministic float32 clumping grid C = 1 + slope*(n/nbar) (type_of_clumping 5 semantics:
    the kernels only see float32 values per cell, clumping_module.F90:18)"""
    return (1.0 + slope * (ndens.astype(np.float64) / avg_dens(zred))).astype(np.float32)


def bubble_state(shape, srcpos, radius_cells, x_in=1.0 - 1e-4, x_out=K.xh_initial):
    """Mid-reionization ionization state: ionized spheres (periodic) of `radius_cells` around
    every source, neutral elsewhere."""
    n3, n2, n1 = shape
    xh = np.full(shape, x_out, dtype=np.float64)
    flat = xh.reshape(-1)
    radius = np.broadcast_to(np.asarray(radius_cells, dtype=np.float64), (len(srcpos),))
    cache = {}
    for (i, j, k), r in zip(np.asarray(srcpos), radius):
        ri = int(np.ceil(r))
        if ri not in cache:
            di = np.arange(-ri, ri + 1, dtype=np.int32)
            dz, dy, dx = np.meshgrid(di, di, di, indexing="ij")
            d2 = (dx * dx + dy * dy + dz * dz).reshape(-1)
            order = np.argsort(d2, kind="stable")
            cache[ri] = (dx.reshape(-1)[order], dy.reshape(-1)[order], dz.reshape(-1)[order], d2[order])
        dx, dy, dz, d2 = cache[ri]
        n = int(np.searchsorted(d2, r * r, side="right"))      # offsets are sorted by distance
        idx = (((k - 1 + dz[:n]) % n3) * n2 + ((j - 1 + dy[:n]) % n2)) * n1 + ((i - 1 + dx[:n]) % n1)
        flat[idx] = x_in
    return xh


def random_state(shape, seed, lo=1e-4, hi=1.0 - 1e-6):
    """log-uniform neutral fraction between 1-hi and 1-lo (test input, not physical)"""
    rng = np.random.default_rng(seed)
    lx = rng.uniform(np.log10(1.0 - hi), np.log10(1.0 - lo), size=shape)
    return 1.0 - 10.0 ** lx
