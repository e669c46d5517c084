#!/usr/bin/env python
"""Benchmark of the B200-native C2-Ray3Dm photo-ionization hot path.

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU restatement of the reference

metric: cell-source ray-trace updates/s (one update = one evolve0D call passing the gate of
evolve_point.F90:128), whole job.  A step is one evolve3D(dt) call (evolve.F90:83): all outer
iterations of {ray-trace every source, reduce the rate grid over ranks, per-cell chemistry}.

Workload (N=1): BASELINE.json configs[2] -- synthetic log-normal density 256^3, 10^4 sources at the
density peaks, clumping grid on, LLS on, mid-reionization bubble state (mean ionized fraction 0.54: spheres of up to
25 cells around the sources) -- the largest configuration that fits one GPU step in seconds.  Every step is one
evolve3D(dt) call from the SAME snapshot (S1), restored on the device before the call, so the time per step is
stationary; the early-reionization state S0 (xh = 2e-4) is measured beside it (key "S0").  With --gpus N the job is
N periodic copies of that volume in one mesh (weak scaling: 256x256x512, 256x512x512, 512^3 for 2, 4, 8 GPUs; N x 10^4
sources, the same density, bubbles and fluxes around every copy of a source, so every GPU has exactly the 1-GPU
updates to do; --scaling strong keeps --mesh and --nsrc for the whole job): every GPU holds the full grids
and traces its round-robin share (master_slave.F90:85), the partial rate grids are summed with ncclAllReduce
(evolve.F90:599-602).  After the timed legs the sampled-source rate grid of the cpu_baseline leg is compared with
the GPU's ("parity_rel_err", must be <= 1e-6).

The reference is Fortran and cannot be built in this image (no Fortran compiler), so the reference arm
and cpu_baseline time the C restatement in oracle/ ("kind": "port") on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "cell-source raytrace updates/s"
UNIT = "updates/s"
B_RT = 28.0  # algorithmic HBM bytes per ray-trace update: ndens 4 + xh_av 8 + phih RMW 16
# dram__bytes_read.sum + dram__bytes_write.sum of one raytrace_kernel launch on this workload divided by
# the updates of that launch (ncu --set full capture, profiles/ncu_raytrace_r2_summary.txt)
NCU_DRAM_BYTES_PER_UPDATE = 20.4
# the same capture (profiles/ncu_raytrace_r2_summary.txt): measured counters, not static instruction counts
NCU_SOURCE = "profiles/ncu_raytrace_r2_summary.txt"
NCU_DRAM_PCT = 35.8         # gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
NCU_FP64_PCT = 45.5         # sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
NCU_ISSUE_PCT = 54.3        # smsp__issue_active.avg.pct_of_peak_sustained_active
NCU_INSTR_PER_UPDATE = 4.39 # smsp__inst_executed.sum / updates of the launch
YEAR = 3.15576e7


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mesh", type=int, default=256)
    ap.add_argument("--nsrc", type=int, default=10000, help="sources per GPU")
    ap.add_argument("--dt-myr", type=float, default=0.5)
    ap.add_argument("--bubble", type=float, default=25.0, help="radius (cells) of the brightest source's bubble")
    ap.add_argument("--cpu-sample", type=int, default=0, help="sources in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-s0", action="store_true", help="skip the early-reionization (S0) extra measurement")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --nsrc sources per GPU (default); strong: --nsrc sources in total")
    return ap.parse_args()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def tiling(world):
    """(tx, ty, tz): how many copies of the base volume along x, y, z for `world` GPUs -- the most even factorisation,
    larger factors on the slower axes (1,1,2), (1,2,2), (2,2,2) for 2, 4, 8"""
    f = [1, 1, 1]
    n, p = world, 2
    primes = []
    while n > 1:
        while n % p == 0:
            primes.append(p)
            n //= p
        p += 1
    for q in sorted(primes, reverse=True):
        i = min(range(3), key=lambda k: (f[k], -k))
        f[i] *= q
    return tuple(sorted(f))


def job_shape(args, world):
    """(mesh, sources in total, bubble radius, tiles) of the job on `world` GPUs; mesh is an int (cubic) or (n1,n2,n3).
    weak scaling: the --mesh^3 / --nsrc volume repeated `world` times in one periodic mesh (tiles along x, y, z);
    strong scaling: --mesh and --nsrc describe the whole job."""
    if args.scaling == "strong" or world == 1:
        return args.mesh, args.nsrc, args.bubble, (1, 1, 1)
    t = tiling(world)
    return (args.mesh * t[0], args.mesh * t[1], args.mesh * t[2]), args.nsrc * world, args.bubble, t


def ncells(mesh):
    return int(mesh) ** 3 if np.isscalar(mesh) else int(mesh[0]) * int(mesh[1]) * int(mesh[2])


def mesh_name(mesh):
    return "%d^3" % mesh if np.isscalar(mesh) else "%dx%dx%d" % tuple(mesh)


def _build_workload(mesh, nsrc_total, bubble, tiles=(1, 1, 1)):
    from c2ray3dm_b200 import synthetic as syn
    zred = 9.0
    ntile = tiles[0] * tiles[1] * tiles[2]
    base = mesh if np.isscalar(mesh) else int(mesh[0]) // tiles[0]
    nsrc = nsrc_total // ntile
    seed = 20240607 if base == 256 else (20240608 if base == 512 else 20240600 + base)
    nd = syn.lognormal_density(base, zred, seed)
    pos, nf = syn.sources_at_density_peaks(nd, nsrc, 1e7)
    radius = bubble * (nf / nf.max()) ** (1.0 / 3.0)
    xh = syn.bubble_state(nd.shape, pos, radius)
    dr, vol = syn.proper_geometry(base, zred)
    clump = syn.clumping_from_density(nd, zred)
    if ntile > 1:
        # N periodic copies of the volume; the copies of a source are consecutive in the list, so the static
        # round-robin of the first pass (master_slave.F90:85) gives every rank one copy of every source
        reps = (tiles[2], tiles[1], tiles[0])   # arrays are [z][y][x]
        nd, xh, clump = np.tile(nd, reps), np.tile(xh, reps), np.tile(clump, reps)
        offs = np.array([(ix * base, iy * base, iz * base) for iz in range(tiles[2]) for iy in range(tiles[1])
                         for ix in range(tiles[0])], dtype=pos.dtype)
        pos = (pos[:, None, :] + offs[None, :, :]).reshape(-1, 3)
        nf = np.repeat(nf, ntile)
    return dict(zred=zred, ndens=np.ascontiguousarray(nd), srcpos=np.ascontiguousarray(pos), normflux=nf,
                xh=np.ascontiguousarray(xh), dr=dr, vol=vol, clumping=np.ascontiguousarray(clump),
                coldensh_LLS=syn.lls_coldens(dr[0], zred))


def build_workload(mesh, nsrc_total, bubble, tiles=(1, 1, 1)):
    """inputs of configs[2]/[3], deterministic; `bubble` = radius of the brightest source's sphere.
    Under torchrun the local rank 0 builds them once and the other ranks of the node read them from /dev/shm."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 or not os.path.isdir("/dev/shm"):
        return _build_workload(mesh, nsrc_total, bubble, tiles)
    tag = "/dev/shm/c2b_workload_%s_%s_%d_%.4f_%s" % (os.environ.get("MASTER_PORT", "0"), mesh_name(mesh).replace("^", "c"),
                                                      nsrc_total, bubble, os.environ.get("TORCHELASTIC_RUN_ID", "run"))
    keys = ("ndens", "srcpos", "normflux", "xh", "dr", "clumping")
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        w = _build_workload(mesh, nsrc_total, bubble, tiles)
        for k in keys:
            np.save(tag + "_" + k + ".npy", w[k])
        with open(tag + ".json.tmp", "w") as f:
            json.dump({"zred": w["zred"], "vol": w["vol"], "coldensh_LLS": w["coldensh_LLS"]}, f)
        os.rename(tag + ".json.tmp", tag + ".json")
        w["_shm_tag"] = tag
        return w
    t0 = time.time()
    while not os.path.exists(tag + ".json"):
        time.sleep(0.5)
        if time.time() - t0 > 1800:
            raise SystemExit("bench.py: timed out waiting for local rank 0 to build the workload")
    with open(tag + ".json") as f:
        w = json.load(f)
    for k in keys:
        w[k] = np.load(tag + "_" + k + ".npy")
    return w


def remove_shared_workload(w):
    """the local rank 0 removes the workload files it shared through /dev/shm"""
    if w.get("_shm_tag"):
        import glob
        for f in glob.glob(w["_shm_tag"] + "*"):
            try:
                os.remove(f)
            except OSError:
                pass


class ClockSampler(threading.Thread):
    """samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [x.strip() for x in out.stdout.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_sample(w, mesh, nsample, threads, keep=False, in_source=False):
    """times the C restatement (oracle/) on a bounded sample: one pass over every k-th source + one
    per-cell pass, on `threads` host threads (sources dealt to threads, private rate grids summed: the
    reference's MPI picture).  Returns (updates/s, description, seconds, updates[, source selection, rate grid])."""
    from oracle import oracle as O
    ns = len(w["normflux"])
    stride = max(1, ns // max(1, nsample))
    sel = np.arange(0, ns, stride)[:nsample]
    o = O.Oracle(mesh)
    o.set_density(w["ndens"])
    o.set_geometry(w["dr"], w["vol"])
    o.set_clumping(5, 1.0, w["clumping"])
    o.set_lls(True, 1, w["coldensh_LLS"], None, 0.0)
    o.set_sources(w["srcpos"][sel], w["normflux"][sel], 1e48)
    o.set_xh(w["xh"])
    o.set_threads(threads)
    o.set_omp_in_source(in_source)
    o.xh_av[...] = w["xh"]
    o.xh_intermed[...] = w["xh"]
    o.state_before()
    o.set_rates_to_zero()
    t0 = time.perf_counter()
    r = o.pass_all_sources()
    t1 = time.perf_counter()
    phih = o.phih.copy() if keep else None
    o.global_pass(0.5e6 * YEAR, r.photon_loss_all)
    t2 = time.perf_counter()
    desc = ("%d of %d sources (every %dth, file order) of the %s workload: 1 pass_all_sources (%.2fs) + "
            "1 global_pass (%.2fs), C restatement, %d threads, mode: %s" % (
                len(sel), ns, stride, mesh_name(mesh), t1 - t0, t2 - t1, threads,
                "omp-in-source (all threads inside one source: 6 axes / 12 planes / 8 octants, evolve_source.F90:141-186)"
                if in_source else "source-parallel (one source per thread, private rate grids summed: do_grid_static + "
                                  "MPI_ALLREDUCE)"))
    if keep:
        return r.updates / (t1 - t0), desc, t2 - t0, r.updates, sel, phih
    return r.updates / (t1 - t0), desc, t2 - t0, r.updates


def cpu_threads(mesh):
    """host threads for the CPU restatement: all cores, unless their private grids (column density + rate grid per
    thread, 16 B per cell) would take more than a quarter of the host memory or 64 GB"""
    cores = os.cpu_count() or 1
    try:
        import psutil
        budget = min(0.25 * psutil.virtual_memory().total, 64e9)
    except Exception:
        budget = 16e9
    return max(1, min(cores, int(budget / (16.0 * ncells(mesh)))))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    mesh_job, nsrc_total, bubble, tiles = job_shape(args, args.gpus)
    cores = cpu_threads(mesh_job)
    w = build_workload(mesh_job, nsrc_total, bubble, tiles)
    nsample = args.cpu_sample or max(cores * 4, 64)
    rates, secs, upd = [], [], []
    desc = ""
    # size the sample towards ~10-20 s of CPU work per step (the whole run stays within minutes)
    v, desc, s, u = cpu_sample(w, mesh_job, nsample, cores)
    if not args.cpu_sample:
        nsample = int(min(len(w["normflux"]), max(8, nsample * 12.0 / max(s, 1e-3))))
    for i in range(args.warmup + args.steps):
        v, desc, s, u = cpu_sample(w, mesh_job, nsample, cores)
        if i >= args.warmup:
            rates.append(v)
            secs.append(s)
            upd.append(u)
    value = float(np.sum(upd) / np.sum([u / r for u, r in zip(upd, rates)]))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference is Fortran and cannot be built in this image (no Fortran compiler): this is the C "
                    "restatement in oracle/ on the host cores"}
    print(json.dumps(line))
    remove_shared_workload(w)
    return 0


def workload_config(args, world=None):
    world = world or args.gpus
    mesh, total, bubble, tiles = job_shape(args, world)
    weak = world > 1 and args.scaling == "weak"
    return {"workload": "synthetic lognormal density %s, %d sources in total (%s) at density peaks, clumping grid "
                        "(type 5), LLS type 1, z=9, dt=%g Myr; every step = one evolve3D call from the same "
                        "mid-reionization snapshot S1 (bubbles r<=%.1f cells, mean x=0.54), restored on the "
                        "device before the call (BASELINE configs[%d]%s)" % (
                            mesh_name(mesh), total,
                            ("weak scaling: %d periodic copies (%dx%dx%d) of the %d^3 / %d-source volume in one mesh, every "
                             "GPU has the 1-GPU updates to do" % (world, tiles[0], tiles[1], tiles[2], args.mesh, args.nsrc))
                            if weak else ("strong scaling" if world > 1 else "one GPU"), args.dt_myr, bubble,
                            3 if args.mesh == 512 else 2, ", weak-scaled" if weak else ""),
            "mesh": mesh if np.isscalar(mesh) else list(mesh), "sources_total": total,
            "parallelism": "source-sharded x%d" % world,
            "state": "S1 restored every step (stationary)",
            "l2": "grids (tau_cell + twin + phih + twin = %.0f MB) exceed the 126 MB L2" % (32 * ncells(mesh) / 1e6)
            if ncells(mesh) >= 256 ** 3 else "grids fit in L2; L2 flushed between steps by the chemistry pass"}


def run_ours(args):
    import ctypes
    import torch
    import torch.distributed as dist
    from c2ray3dm_b200 import Evolve

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA library is the product (no CPU fallback). "
                         "Use --impl reference for the CPU restatement.")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mesh, nsrc_total, bubble, tiles = job_shape(args, world)
    w = build_workload(mesh, nsrc_total, bubble, tiles)
    nc = ncells(mesh)
    dt = args.dt_myr * 1e6 * YEAR

    e = Evolve(mesh, device=local, rank=rank, nranks=world, type_of_clumping=5, use_LLS=True, type_of_LLS=1)
    if world > 1:
        uid = [Evolve.get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        e.comm_init(uid[0])
    e.rad_ini()
    e.set_geometry(w["dr"], w["vol"])
    e.set_clumping(w["clumping"])
    e.set_LLS(coldensh_LLS=w["coldensh_LLS"])
    e.set_sources(w["srcpos"], w["normflux"])
    # pinned host staging for the end-to-end leg
    nd_pin = torch.from_numpy(w["ndens"].reshape(-1)).pin_memory()
    xh_pin = torch.from_numpy(w["xh"].reshape(-1).copy()).pin_memory()
    e.set_density(nd_pin.numpy())
    e.set_xh(xh_pin.numpy())
    e.save_xh()      # the S1 snapshot every step starts from (device-resident)

    def barrier():
        torch.cuda.synchronize()
        e.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # Every step is one evolve3D(dt) call from the SAME mid-reionization snapshot S1, restored on the device
    # (134 MB device-to-device at 256^3, ~0.1 ms, inside the timed region), so ms_per_step is stationary and
    # comparable across N and rounds.
    tsim = 0.0
    for _ in range(args.warmup):
        e.restore_xh()
        e.evolve3D(tsim, dt)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region 1: state resident in HBM -------------------------------------------------
    barrier()
    t0 = time.perf_counter()
    upd = launches = 0
    ms_rt = ms_chem = ms_ar = ms_dev = 0.0
    niter = 0
    step_ms = []
    for _ in range(args.steps):
        e.restore_xh()
        rep = e.evolve3D(tsim, dt)
        upd += rep.total_updates
        launches += rep.kernel_launches
        ms_rt += rep.ms_raytrace
        ms_chem += rep.ms_chemistry
        ms_ar += rep.ms_allreduce
        ms_dev += rep.ms_total
        step_ms.append(rep.ms_total)
        niter += rep.niter
    barrier()
    wall = maxreduce(time.perf_counter() - t0)
    ms_rt_max = maxreduce(ms_rt)
    # a rank that finishes its traces early waits inside the all-reduce: the smallest value over the ranks is the
    # collective itself, the largest includes the wait for the slowest rank
    ms_ar_min = -maxreduce(-ms_ar)
    ms_ar_max = maxreduce(ms_ar)
    # ---- timed region 2: end to end through the C ABI with host buffers ---------------------------
    # what fortran/evolve_b200.F90 moves per evolve3D call: ndens + xh in, xh + phih_grid out
    e2e = None
    dptr = ctypes.POINTER(ctypes.c_double)
    if not args.no_e2e:
        xh_out = torch.empty(nc, dtype=torch.float64).pin_memory()
        ph_out = torch.empty(nc, dtype=torch.float64).pin_memory()
        barrier()
        t0 = time.perf_counter()
        upd2 = 0
        for _ in range(args.steps):
            e.set_density(nd_pin.numpy())     # what the host re-sends after cosmo_evol (cosmology.F90:186)
            e.set_xh(xh_pin.numpy())          # ionfractions_module.F90:22 (the S1 snapshot again)
            rep = e.evolve3D(tsim, dt)
            e._ck(e.L.c2b_get_xh(e.h, xh_out.numpy().ctypes.data_as(dptr)), "c2b_get_xh")
            e._ck(e.L.c2b_get_phih(e.h, ph_out.numpy().ctypes.data_as(dptr)), "c2b_get_phih")
            upd2 += rep.total_updates
        barrier()
        wall2 = maxreduce(time.perf_counter() - t0)
        e2e = {"value": upd2 / wall2, "unit": UNIT, "h2d_bytes_per_step": 12 * nc,
               "d2h_bytes_per_step": 16 * nc, "ms_per_step": 1e3 * wall2 / args.steps}
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)
    # ---- early-reionization state S0: xh = 2e-4 everywhere, every trace ends in subbox 1-3 ------
    s0 = None
    if not args.no_s0:
        from c2ray3dm_b200 import constants as K
        xh0 = torch.full((nc,), K.xh_initial, dtype=torch.float64).pin_memory()
        u0 = rt0 = 0.0
        t_s0 = 0.0
        for i in range(3):
            e.set_xh(xh0.numpy())
            barrier()
            t0 = time.perf_counter()
            rep = e.evolve3D(tsim, dt)
            barrier()
            if i > 0:   # the first step is the warm-up (it also re-learns the per-source trace lengths)
                t_s0 += maxreduce(time.perf_counter() - t0)
                u0 += rep.total_updates
                rt0 += rep.ms_raytrace
        rt0 = maxreduce(rt0)
        s0 = {"state": "S0: xh = 2e-4 everywhere (every trace ends in subbox 1-3)", "value": u0 / t_s0, "unit": UNIT,
              "ms_per_step": 1e3 * t_s0 / 2, "updates_per_step": u0 / 2,
              "raytrace_updates_per_s_per_gpu": (u0 / world) / (rt0 * 1e-3) if rt0 > 0 else None,
              "raytrace_frac_of_hbm_roofline": (u0 / world) * B_RT / (rt0 * 1e-3) / 1e9 / load_peaks()[0] if rt0 > 0 else None}
        e.set_xh(xh_pin.numpy())
    peak, peak_kind = load_peaks()
    dfma = e.measure_dfma_rate()
    line = None
    if rank == 0:
        # updates of THIS rank's ray-trace kernels over their device time (CUDA events on the launch stream)
        upd_rank = upd / world
        achieved = upd_rank * B_RT / (ms_rt_max * 1e-3) / 1e9 if ms_rt_max > 0 else 0.0
        line = {"metric": METRIC, "value": upd / wall, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, world),
                "seconds_per_evolve3D_step": wall / args.steps,
                "step_ms": step_ms,
                "updates_per_step": upd / args.steps, "outer_iterations_per_step": niter / args.steps,
                "gpu_launches": int(launches),
                "phase_ms_per_step": {"raytrace": ms_rt / args.steps, "raytrace_max_over_ranks": ms_rt_max / args.steps,
                                      "allreduce": ms_ar / args.steps, "allreduce_min_over_ranks": ms_ar_min / args.steps,
                                      "allreduce_max_over_ranks": ms_ar_max / args.steps,
                                      "chemistry": ms_chem / args.steps, "device_total": ms_dev / args.steps},
                "roofline": {"kernel": "raytrace_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak,
                             "traffic": NCU_DRAM_BYTES_PER_UPDATE * upd_rank / max(1, niter) if (args.mesh == 256 and args.bubble == 25.0) else None,
                             "traffic_note": "bytes per launch = %.1f B/update (ncu dram bytes of one launch of this "
                                             "workload, profiles/) x updates per launch" % NCU_DRAM_BYTES_PER_UPDATE,
                             "algorithmic_bytes_per_launch": B_RT * upd_rank / max(1, niter),
                             "peak_source": "%s (MEASURED_PEAKS.json hbm_gbs)" % peak_kind,
                             "algorithmic_bytes_per_update": B_RT,
                             "note": "the kernel is issue/latency-bound, not bandwidth-bound (profiles/): DRAM "
                                     "throughput and FP64-pipe activity measured by ncu are in `ncu`"},
                "ncu": {"source": NCU_SOURCE, "dram_throughput_pct": NCU_DRAM_PCT, "fp64_pipe_active_pct": NCU_FP64_PCT,
                        "issue_active_pct": NCU_ISSUE_PCT, "warp_instr_per_update": NCU_INSTR_PER_UPDATE,
                        "dfma_per_s_probe": dfma},
                "clocks": sampler.summary(),
                "S0": s0,
                "e2e": e2e}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = cpu_threads(mesh)
        nsample = args.cpu_sample or max(cores * 4, 64)
        v, desc, s, u, sel, ph_cpu = cpu_sample(w, mesh, nsample, cores, keep=True)
        if s < 5:  # scale the sample towards ~10-30 s of CPU work
            nsample = min(len(w["normflux"]), int(nsample * 15 / max(s, 1e-3)))
            v, desc, s, u, sel, ph_cpu = cpu_sample(w, mesh, nsample, cores, keep=True)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
        # the reference's other parallel mode: OpenMP inside each source, at most 12-way
        v2, desc2, s2, u2 = cpu_sample(w, mesh, max(8, len(sel) // 6), cores, in_source=True)
        line["cpu_baseline"]["omp_in_source"] = {"value": v2, "unit": UNIT, "cores": cores, "sample": desc2}
        # parity of the timed workload: the same sampled sources, one pass from the S1 snapshot, GPU vs CPU
        e.set_sources(w["srcpos"][sel], w["normflux"][sel])
        e.set_xh(xh_pin.numpy())
        e.begin_step()
        g = e.pass_all_sources()
        ph_gpu = e.phih_grid.reshape(-1)
        ph_cpu = ph_cpu.reshape(-1)
        nz = ph_cpu != 0
        err = float(np.max(np.abs(ph_gpu[nz] - ph_cpu[nz]) / ph_cpu[nz]))
        line["parity_rel_err"] = err
        line["parity"] = {"what": "phih_grid of the cpu_baseline sample (%d sources, one pass from S1): max relative "
                                  "difference GPU vs CPU restatement; update counts equal: %s; cells with a rate equal: %s"
                                  % (len(sel), g.updates == u, bool(np.array_equal(ph_gpu != 0, nz))),
                          "tolerance": 1e-6, "ok": bool(err <= 1e-6 and g.updates == u and np.array_equal(ph_gpu != 0, nz))}
        if not line["parity"]["ok"]:
            print(json.dumps(line))
            raise SystemExit("bench.py: the timed workload does not match the CPU restatement (parity_rel_err %.3e)" % err)
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    e.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    remove_shared_workload(w)
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
