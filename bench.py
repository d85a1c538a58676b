#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Benchmark of the tasmania stencil hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

Workload (BASELINE.json): dry isentropic model, flow over an isolated Gaussian mountain,
relaxed lateral boundaries (nb=3, nr=6), RK3WS + fifth-order upwind, Rayleigh damping
(depth 15, max 5e-4), dt = 5 s, fp64.  Default size = the configuration the metric's target is
quoted on: 1024x1024x64 per GPU (config 5; weak scaling over a 2-D decomposition), which also
keeps every field (0.56 GB) far larger than the 126 MB L2, so no L2 flush is needed between
timed steps.  ``--workload c2`` runs the 161x161x60 case of config 2 (L2-resident,
launch-latency regime); ``--workload c4`` runs configs[3], the fourth-order diffusion dwarf at
4096x4096x64 with periodic boundaries (its own loop and JSON line: ``run_c4``).

One *step* = ``dycore.update_topography`` + one full RK3WS step (3 fused stages) + the
``IsentropicDiagnostics`` refresh of p / exn / mtg / h.
Metric: grid-point updates per second, Mpts*steps/s = nx*ny*nz*steps*N / seconds / 1e6.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from datetime import datetime, timedelta

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

P, EXN, H = ("air_pressure_on_interface_levels", "exner_function_on_interface_levels",
             "height_on_interface_levels")

WORKLOADS = {
    # name: (nx, ny, nz) per GPU
    "c5": (1024, 1024, 64),
    "c5halo": (1028, 1024, 64),  # local grid of a rank of the 2x1 decomposition (timing experiments)
    "c2": (161, 161, 60),
    "small": (256, 256, 64),
    # configs[3]: the fourth-order horizontal-diffusion dwarf on a doubly periodic grid (a different
    # loop: run_c4 below); c4small is the same loop at a size for quick checks
    "c4": (4096, 4096, 64),
    "c4small": (512, 512, 64),
}
# algorithmic HBM bytes per grid point (SURVEY.md section 8d / BASELINE.md section 4):
# dry RK3WS step 3 x 112 B + diagnostics refresh 40 B
BYTES_PER_POINT_STEP = 336 + 40
BYTES_PER_POINT_STAGE = 112


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ workload set-up (host)
def mountain_case(nx, ny, nz):
    """BASELINE config 2/5 initial condition on an (nx, ny, nz) grid; returns (grid, numpy state).
    The grid spacing of config 2 is kept (dx = dy = 352 km / 160 = 2.2 km, so that dt = 5 s stays
    inside the stability limit of the gravity waves): the domain is +-1.1 (n - 1) km wide, i.e.
    +-176 km at 161 points and +-1125 km at 1024.  Columns of the initial state are horizontally
    uniform (the mountain has not grown yet), so one column is built with the reference formulas
    and broadcast."""
    from tasmania_b200.grid import Grid, Topography, gaussian_profile, isentropic_state_from_brunt_vaisala

    hx, hy = 1.1 * (nx - 1), 1.1 * (ny - 1)
    x = np.linspace(-hx, hx, nx)
    y = np.linspace(-hy, hy, ny)
    topo = Topography(gaussian_profile(x, y, 500.0, 50.0, 50.0), timedelta(seconds=1800))
    grid = Grid((-hx, hx), nx, (-hy, hy), ny, (400.0, 280.0), nz, units_to_m=1e3,
                topography=topo)
    small = Grid((-176.0, 176.0), 3, (-176.0, 176.0), 3, (400.0, 280.0), nz, units_to_m=1e3)
    col = isentropic_state_from_brunt_vaisala(small, 22.5, 0.0, 0.015)
    state = {}
    for name, a in col.items():
        full = np.zeros((nx + 1, ny + 1, nz + 1))
        mi = nx + 1 if "at_u_locations" in name else nx
        mj = ny + 1 if "at_v_locations" in name else ny
        full[:mi, :mj, :] = a[1, 1, :][None, None, :]
        state[name] = full
    return grid, state


class DryRun:
    """The timed loop on b200 storages (tasmania_b200.isentropic_dry.IsentropicDryRun on the
    benchmark's initial condition); ``graph=True`` replays the step as CUDA graphs."""

    def __init__(self, nx, ny, nz, device_index=0, graph=False):
        import torch

        import tasmania_b200 as tb
        from tasmania_b200.isentropic import MTG, S
        from tasmania_b200.isentropic_dry import IsentropicDryRun

        self.tb, self.torch = tb, torch
        self.S, self.MTG = S, MTG
        self.nx, self.ny, self.nz = nx, ny, nz
        self.grid, np_state = mountain_case(nx, ny, nz)
        self.np_state = np_state
        self.run = IsentropicDryRun(self.grid, np_state, timedelta(seconds=5))
        self.out_names = self.run.out_names
        self.names = (S,) + tuple(n for n in self.out_names if n != S) + (MTG,)
        self.pt, self.dt, self.hb, self.dyc, self.diag = (self.run.pt, self.run.dt, self.run.hb,
                                                          self.run.dyc, self.run.diag)
        assert self.dyc._fused
        self.loop = None
        if graph:
            from tasmania_b200.graphs import GraphedLoop

            self.loop = GraphedLoop(self.run)

    @property
    def state(self):
        return self.run.state

    def step(self):
        if self.loop is not None:
            self.loop.step()
        else:
            self.run.step()


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU baseline (oracle port)
def cpu_baseline(seconds_budget=20.0, grid=(161, 161, 60)):
    """Time the numpy oracle (a port of the reference's numpy backend; the reference itself is
    Python and cannot travel to the GPU box) on a bounded sample of the same workload: full
    RK3WS steps of the same physics on a 161x161x60 grid (config 2 size), as many as fit the
    budget (at least 1).  numpy element-wise code is single-threaded."""
    from oracle import boundary as ob
    from oracle import isentropic as oi

    nx, ny, nz = grid
    g, np_state = mountain_case(nx, ny, nz)
    pt = float(np_state[P][0, 0, 0])
    ogrid = oi.Grid(nx, ny, nz, g.dx, g.dy, g.dz, g.z_on_interface_levels, g.z)
    ohb = ob.Relaxed(nx, ny, nz, 3, 6)
    state = {n: v.copy() for n, v in np_state.items()}
    state["time"] = datetime(2000, 1, 1)
    ohb.reference_state = {n: v.copy() for n, v in np_state.items()}
    topo = g.topography
    dyc = oi.IsentropicDycore(ogrid, ohb, lambda: topo.profile, scheme="rk3ws_si",
                              flux="fifth_order_upwind", pt=pt, eps=0.5, damp=True, damp_depth=15,
                              damp_max=5e-4)
    dt = timedelta(seconds=5)
    t0 = time.perf_counter()
    steps = 0
    while True:
        topo.update((steps + 1) * dt)
        out = dyc(state, {}, dt)
        new = {n: out[n].copy() for n in (oi.S, oi.SU, oi.U, oi.SV, oi.V)}
        new["time"] = out["time"]
        for n in (P, EXN, H, oi.MTG):
            new[n] = state[n]
        oi.refresh_diagnostics(ogrid, topo.profile, new[oi.S], pt, new[P], new[EXN], new[oi.MTG], new[H])
        state = new
        steps += 1
        if time.perf_counter() - t0 > seconds_budget or steps >= 50:
            break
    el = time.perf_counter() - t0
    return {
        "value": nx * ny * nz * steps / el / 1e6,
        "unit": "Mpts*steps/s",
        "cores": 1,
        "kind": "port",
        "sample": f"{steps} RK3WS steps (+diagnostics refresh) of the same dry isentropic workload on "
                  f"{nx}x{ny}x{nz}, numpy oracle, 1 thread of {os.cpu_count()} host cores, "
                  f"{el:.1f} s; the reference's gt4py CPU backends are not installable offline",
    }, steps, el


def reference_root():
    """Where the UNMODIFIED reference sources are, if they are around: $TASMANIA_REFERENCE, or the
    copy staged by baseline/stage_reference.sh under the git-ignored baseline/_ref (which travels
    to the GPU box with the snapshot).  /root/reference itself is never read here."""
    for cand in (os.environ.get("TASMANIA_REFERENCE"), os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "src", "tasmania")):
            return cand
    return None


def cpu_baseline_reference(seconds_budget=20.0, grid=(161, 161, 60)):
    """The reference ITSELF on its numpy backend (kind "reference"): its own Domain, Relaxed
    boundary, state builder, RK3WSSI prognostic, Rayleigh damper, HorizontalVelocity,
    IsentropicDiagnostics and ``IsentropicDynamicalCore.stage_array_call_dry``, imported in place
    from the staged sources and chained like framework/dycore.py does (tests/ref_dycore_steps.py;
    only the sympl wrappers, which are not installable offline, are replaced by that driver).
    Same workload as ``cpu_baseline``: RK3WS + fifth-order upwind steps + diagnostics refresh."""
    os.environ["TASMANIA_REFERENCE"] = reference_root()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_dycore_steps as rds

    nx, ny, nz = grid
    clock = {"budget": seconds_budget}
    rds.run("numpy", 50, nx, ny, nz, clock=clock, damp_depth=15, topo_seconds=1800)
    steps, el = clock["steps"], clock["seconds"]
    return {
        "value": nx * ny * nz * steps / el / 1e6,
        "unit": "Mpts*steps/s",
        "cores": 1,
        "kind": "reference",
        "sample": f"{steps} RK3WS steps (+diagnostics refresh) of the same dry isentropic workload on "
                  f"{nx}x{ny}x{nz} by the UNMODIFIED reference (stubbiali/tasmania, numpy backend, its own "
                  f"stage_array_call_dry; sources staged under baseline/_ref), 1 thread of "
                  f"{os.cpu_count()} host cores, {el:.1f} s; the reference's gt4py CPU backends are "
                  f"not installable offline",
    }, steps, el


def _replica(job):
    """One replica of the bounded sample (runs in a spawned worker process)."""
    seconds_budget, grid = job[:2]
    fn = cpu_baseline_reference if (len(job) > 2 and job[2] == "reference") else cpu_baseline
    base, steps, el = fn(seconds_budget=seconds_budget, grid=grid)
    return base["value"], steps, el


def cpu_baseline_one(seconds_budget, grid, kind):
    """One timed replica in a fresh process (the reference's loader installs import stubs for the
    packages it cannot have offline: kept out of the benchmark process)."""
    import multiprocessing as mp

    with mp.get_context("spawn").Pool(1) as pool:
        fn = cpu_baseline_reference if kind == "reference" else cpu_baseline
        return pool.apply(fn, (seconds_budget, grid))


def cpu_baseline_all_cores(seconds_budget, grid, kind="port"):
    """The reference has no threading and no domain decomposition (SURVEY.md header): one Python
    thread drives single-threaded numpy element-wise code.  The only way its code uses every host
    core is one independent replica of the workload per core, which is what is timed here: the
    aggregate is the throughput the box's host cores deliver with the reference's algorithm
    (memory-bandwidth contention between the replicas included)."""
    import multiprocessing as mp

    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    cores = min(cores, 32)  # bounds host memory (a replica holds ~0.5 GB of fields and temporaries)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_replica, [(seconds_budget, grid, kind)] * cores, chunksize=1)
    return cores, res


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the numpy oracle port
    (the reference is Python + absent on the GPU box; its numpy backend is what the oracle
    restates and is pinned against), on all host cores (one replica per core, see
    ``cpu_baseline_all_cores``).  Each 'step' is one RK3WS step on the bounded sample grid."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload.startswith("c4"):
        return run_reference_c4(args)
    nx, ny, nz = WORKLOADS["c2"]
    pts = nx * ny * nz
    budget = min(60.0, 4.0 * max(1, args.steps))
    kind = "reference" if reference_root() is not None else "port"
    one, _, _ = cpu_baseline_one(min(10.0, budget), (nx, ny, nz), kind)
    cores, res = cpu_baseline_all_cores(budget, (nx, ny, nz), kind)
    val = sum(r[0] for r in res)
    steps = sum(r[1] for r in res)
    el = max(r[2] for r in res)
    base = {
        "value": val, "unit": "Mpts*steps/s", "cores": cores, "kind": kind,
        "single_thread_value": one["value"],
        "sample": f"{cores} independent replicas (one per host core; the reference itself is "
                  f"single-threaded) of the same dry isentropic workload on {nx}x{ny}x{nz}, "
                  + ("the UNMODIFIED reference on its numpy backend (sources staged under baseline/_ref), "
                     if kind == "reference" else "numpy oracle (port of the reference's numpy backend), ") +
                  f"{steps} RK3WS steps (+diagnostics refresh) in total in {el:.1f} s; one "
                  f"replica alone: {one['value']:.2f} Mpts*steps/s; the reference's gt4py CPU "
                  f"backends are not installable offline",
    }
    line = {
        "impl": "reference", "metric": "grid-point updates/sec", "value": val,
        "unit": "Mpts*steps/s", "n_gpus": args.gpus, "steps": steps, "warmup": 0,
        "ms_per_step": pts / (val * 1e6) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name("c2").replace(" per GPU", "") + " -- the bounded sample of the "
                               "headline workload (same physics and schemes, config 2's grid; the "
                               f"b200 arm runs {'x'.join(str(v) for v in WORKLOADS[args.workload])} per GPU)",
                   "sample_grid": [nx, ny, nz], "replicas": cores},
        "cpu_baseline": base,
        "e2e": {"value": val, "unit": "Mpts*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_reference_c4(args):
    """--impl reference --workload c4: the numpy oracle port of the diffusion-dwarf loop (one
    thread: numpy element-wise code) on a 512x512x64 sample of the same workload."""
    from oracle import boundary as ob
    from oracle import dwarfs

    nx, ny, nz, nb, dt = 512, 512, 64, 2, 0.05
    hb = ob.Periodic(nx, ny, nz, nb)
    phi = hb.get_numerical_field(np.random.default_rng(20261018).standard_normal((nx, ny, nz)))
    gamma = np.zeros(phi.shape)
    gamma[...] = dwarfs.vertical_profile(0.5, 1.0, 15, nz)[None, None, :]
    tnd = np.zeros(phi.shape)
    budget = min(60.0, 4.0 * max(1, args.steps))
    t0, steps = time.perf_counter(), 0
    while True:
        dwarfs.diffusion(4, phi, gamma, tnd, 1.0, 1.0, True, (nb, nb, 0), (nx, ny, nz))
        phi = phi + dt * tnd
        hb.enforce_field(phi)
        steps += 1
        if time.perf_counter() - t0 > budget or steps >= 200:
            break
    el = time.perf_counter() - t0
    val = nx * ny * nz * steps / el / 1e6
    base = {"value": val, "unit": "Mpts*steps/s", "cores": 1, "kind": "port",
            "sample": f"{steps} applications (diffusion + update + periodic halo) on {nx}x{ny}x{nz}, "
                      f"numpy oracle, 1 thread, {el:.1f} s"}
    print(json.dumps({
        "impl": "reference", "metric": "grid-point updates/sec", "value": val, "unit": "Mpts*steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": 0, "ms_per_step": el / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "fourth-order horizontal diffusion dwarf, periodic BC, fp64",
                   "sample_grid": [nx, ny, nz]},
        "cpu_baseline": base,
        "e2e": {"value": val, "unit": "Mpts*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(key):
    nx, ny, nz = WORKLOADS[key]
    return (f"dry isentropic, gaussian mountain, relaxed BC nb=3 nr=6, RK3WS + fifth_order_upwind, "
            f"{nx}x{ny}x{nz} per GPU, fp64")


# ------------------------------------------------------------------ our arm
def run_b200(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the b200 backend has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    nx, ny, nz = WORKLOADS[args.workload]
    if distributed:
        from tasmania_b200.distributed import DecomposedDryRun

        run = DecomposedDryRun(nx, ny, nz, rank, world, overlap=args.overlap)
    else:
        run = DryRun(nx, ny, nz, local_rank, graph=args.graph)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        run.step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    from tasmania_b200 import lib as tblib

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = tblib.launch_count()
    replayed0 = run.loop.replayed_launches if getattr(run, "loop", None) is not None else 0
    ev0.record()
    for _ in range(args.steps):
        run.step()
    ev1.record()
    barrier()
    launches = tblib.launch_count() - launches0  # kernels of libtasmania_b200.so, this rank
    if getattr(run, "loop", None) is not None:  # launches replayed from CUDA graphs are ours too
        launches += run.loop.replayed_launches - replayed0
    ms = ev0.elapsed_time(ev1)
    if os.environ.get("TB200_DEBUG_RANKS"):
        print(f"[rank {rank}] {ms / args.steps:.3f} ms/step on its own clock", file=sys.stderr, flush=True)
    if distributed:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    if distributed and run.transport == "p2p":
        run.sub.halo.check()  # raises if a pull ever gave up waiting for a neighbour

    # the timed steps must have been a healthy simulation: a state that blew up (NaN / inf) would
    # time special-case arithmetic, not the workload
    for name in run.out_names:
        if not bool(torch.isfinite(run.state[name].t).all()):
            raise RuntimeError(f"bench: {name} is not finite after {args.warmup + args.steps} steps")

    pts = nx * ny * nz
    value = pts * args.steps * world / (ms * 1e-3) / 1e6
    decomposition = getattr(run, "decomposition", "1x1")
    halo_mode = ("overlapped with the interior blocks of the momentum kernel"
                 if getattr(run, "overlap", None) is not None else ("after each stage" if distributed else "none"))
    if distributed:
        halo_mode += {"p2p": "; NVLink peer stores into the neighbours' receive buffers (tb200_halo_push / "
                             "tb200_halo_pull over CUDA IPC), no NCCL on the path, "
                             f"{getattr(run.sub.halo, 'phases', 2)} phase(s) per exchange",
                      "nccl": "; torch.distributed point-to-point messages (NCCL)"}[run.transport]
    api = ("tasmania_b200.distributed.DecomposedDryRun.step" if distributed else
           "tasmania_b200.isentropic_dry.IsentropicDryRun.step") + \
        " -> tasmania_b200.isentropic.IsentropicDynamicalCore.__call__ (mirror of the reference class; " \
        "fused stage tb200_isentropic_stage_dry, lazy velocities)"
    halo_bytes = run.sub.halo.bytes_per_exchange if distributed else None

    # ---- halo exchange share (device time of pack + send/recv + unpack + seam fix-up per stage)
    exchange_ms = None
    if distributed:
        if run.overlap is not None:
            run.overlap.events = []
        else:
            run.exchange_events = []
        for _ in range(2):
            run.step()
        torch.cuda.synchronize()
        exchange_ms = run.exchange_ms()
        run.exchange_events = None
        if run.overlap is not None:
            run.overlap.events = None
    # ---- roofline of the dominant kernel (stage_m: the momentum step), timed live
    roof = kernel_roofline(run, args) if (not distributed or os.environ.get("TB200_DEBUG_RANKS")) else None
    if distributed and roof is not None:
        print(f"[rank {rank}] kernels " + str({k.split()[0]: round(v["ms_per_launch"], 3)
                                                for k, v in roof["kernels"].items()}), file=sys.stderr, flush=True)
        roof = None
    # ---- end to end through the public API with host buffers
    e2e = end_to_end(run, args, world, barrier, distributed)
    aux = None
    if not distributed and not args.no_aux and args.workload == "c5":
        del run  # release the headline's fields first
        torch.cuda.empty_cache()
        aux = aux_measurements(args)
        run = None

    if rank == 0:
        base = None
        if world == 1 and not args.no_cpu_baseline:
            base, _, _ = cpu_baseline_one(20.0, WORKLOADS["c2"],
                                          "reference" if reference_root() is not None else "port")
        line = {
            "metric": "grid-point updates/sec", "value": value, "unit": "Mpts*steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "l2": "inputs larger than L2"
                       if pts * 8 > 126e6 else "L2-resident grid (no flush: launch-latency regime)",
                       "decomposition": decomposition,
                       "halo_exchange": halo_mode, "api": api},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches * world,
            "hbm_frac_step": BYTES_PER_POINT_STEP * pts / (ms / args.steps * 1e-3) / 1e9
            / measured_peak_gbs()[0],
        }
        if roof is not None:
            line["roofline"] = roof
        if aux is not None:
            line["aux"] = aux
        if exchange_ms is not None:
            line["halo_exchange_ms_per_stage"] = exchange_ms
            line["halo_exchange_bytes_per_stage"] = halo_bytes
        if base is not None:
            line["cpu_baseline"] = base
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


# ------------------------------------------------------------------ configs[3]: diffusion dwarf
def run_c4(args):
    """``--workload c4``: BASELINE.json configs[3], the pure-bandwidth stencil.  One step = one
    application of the fourth-order diffusion (16 B/pt), phi <- phi + dt * tnd (24 B/pt) and the
    periodic halo refresh (tasmania_b200.diffusion_dwarf.DiffusionDwarfRun); the roofline object
    is the diffusion kernel timed alone.  Single GPU (under torchrun every rank runs its own
    replica: the periodic wrap is not decomposed)."""
    import torch

    from tasmania_b200 import lib as tblib
    from tasmania_b200.diffusion_dwarf import DiffusionDwarfRun

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the b200 backend has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    nx, ny, nz = WORKLOADS[args.workload]
    run = DiffusionDwarfRun(nx, ny, nz)
    for _ in range(args.warmup):
        run.step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    launches0 = tblib.launch_count()
    ev[0].record()
    for _ in range(args.steps):
        run.step()
    ev[1].record()
    torch.cuda.synchronize()
    launches = tblib.launch_count() - launches0
    ms = ev[0].elapsed_time(ev[1])
    clocks = sampler.stop() if rank == 0 else None
    if not bool(torch.isfinite(run.phi.t).all()):
        raise RuntimeError("bench: phi is not finite")
    # the diffusion kernel alone (phi and tnd are 8.6 GB each at c4: far larger than L2)
    for _ in range(2):
        run.diffusion(run.phi, run.tnd)
    ev[2].record()
    for _ in range(args.steps):
        run.diffusion(run.phi, run.tnd)
    ev[3].record()
    torch.cuda.synchronize()
    k_ms = ev[2].elapsed_time(ev[3]) / args.steps
    pts = nx * ny * nz
    peak, how = measured_peak_gbs()
    gbs = 16 * pts / (k_ms * 1e-3) / 1e9
    if rank == 0:
        print(json.dumps({
            "metric": "grid-point updates/sec", "value": pts * args.steps / (ms * 1e-3) / 1e6,
            "unit": "Mpts*steps/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"fourth-order horizontal diffusion dwarf, periodic BC, {nx}x{ny}x{nz}, "
                                   "fp64; step = diffusion + phi update + halo refresh",
                       "l2": "inputs larger than L2" if pts * 8 > 126e6 else "L2-resident grid"},
            "clocks": clocks, "e2e": None, "gpu_launches": launches,
            "hbm_frac_step": (16 + 24) * pts / (ms / args.steps * 1e-3) / 1e9 / peak,
            "roofline": {"bound": "hbm", "kernel": "diffusion (" + {"march": "march_kernel<4>", "tile": "cross_kernel<4>"}.get(
                os.environ.get("TB200_DIFF_IMPL", ""), "march2_kernel<4>") + ")", "achieved": gbs,
                         "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                         "peak_source": how, "ms_per_launch": k_ms,
                         "algorithmic_bytes_per_launch": 16 * pts},
        }))


# Algorithmic HBM bytes per grid point of each kernel of the fused stage, per RK stage (every
# distinct array read once, every output written once).  With lazy velocities
# (tb200_isentropic_stage.derive_uv_in / skip_uv_out) the three stages of a step differ:
#   s-step     stage 0: s (now == int), u, v -> s_pre = 32; stages 1, 2: s_now, s_int, su_int, sv_int -> s_pre = 40
#   scans      s_pre -> mtg_new = 16
#   momentum   stage 0: s, s_pre, mtg_now, mtg_new, u, v, su, sv -> su, sv (+ s in the relaxation band
#              / damping layer only) = 80; stage 1: s_now, s_int, s_pre, mtg_now, mtg_new, su_now, su_int,
#              sv_now, sv_int -> su, sv = 88; stage 2: the same (u, v of the step's final state come from
#              one pass of tb200_velocity_components: s, su, sv -> u, v = 40 B/pt, timed in `other_kernels`)
KERNEL_NAMES = ("s_step (stage_a_kernel)", "column_scan (stage_b_kernel)", "momentum (stage_mv2_kernel)")
KERNEL_BYTES_PER_POINT_LAZY = {0: (32, 16, 80), 1: (40, 16, 88), 2: (40, 16, 88)}
KERNEL_BYTES_PER_POINT_EAGER = {0: (32, 16, 104), 1: (40, 16, 120), 2: (40, 16, 120)}
NCU_KERNEL_KEYS = ("stage_a_kernel", "stage_b_kernel", "stage_mv2_kernel")
# DRAM traffic per launch: read from the ncu --set full capture of THIS round's build of the same
# command (profiles/README.md says which commit), never typed in
NCU_TRAFFIC_CSV = os.path.join(ROOT, "profiles", "r02_c5_ncu_full_raw.csv")


def ncu_traffic(kernel_key, path=NCU_TRAFFIC_CSV):
    """Mean dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels whose name
    contains ``kernel_key`` in an ``ncu --page raw --csv`` export; (None, None) if unavailable."""
    import csv

    if not os.path.exists(path):
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    try:
        with open(path, newline="") as f:
            rows = list(csv.reader(f))
        head, units = rows[0], rows[1]
        kn, rd, wr = head.index("Kernel Name"), head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")
        vals = [float(r[rd].replace(",", "")) * unit[units[rd]] + float(r[wr].replace(",", "")) * unit[units[wr]]
                for r in rows[2:] if len(r) > max(kn, rd, wr) and kernel_key in r[kn]]
    except (ValueError, KeyError, IndexError, OSError):
        return None, None
    if not vals:
        return None, None
    return float(np.mean(vals)), f"profiles/{os.path.basename(path)} ({len(vals)} launches, mean per launch)"


def kernel_roofline(run, args):
    """Roofline of the dominant kernel (the momentum kernel) and of the whole fused RK stage,
    from CUDA events recorded by the library on the launching stream around each kernel
    (tb200_stage_profile) inside a running time loop."""
    import ctypes as C

    import torch

    from tasmania_b200 import lib as tblib

    handle = tblib.load()
    dyc = run.dyc
    orig = dyc._stage_fused
    samples = {}

    def timed(stage, state, timestep, out_state):
        orig(stage, state, timestep, out_state)
        ms = (C.c_double * 3)()
        tblib.check(handle.tb200_stage_profile_read(ms), "tb200_stage_profile_read")
        samples.setdefault(stage, []).append(tuple(ms))

    tblib.check(handle.tb200_stage_profile(1), "tb200_stage_profile")
    dyc._stage_fused = timed
    eager_step = run.run.step if hasattr(run, "run") else run.step  # never through a graph replay
    for _ in range(max(2, min(args.steps, 5))):
        eager_step()
    torch.cuda.synchronize()
    dyc._stage_fused = orig
    tblib.check(handle.tb200_stage_profile(0), "tb200_stage_profile")
    per_stage = {st: np.mean(np.array(v), axis=0) for st, v in samples.items()}  # ms: s-step, scan, momentum
    nst = len(per_stage)
    bpp = KERNEL_BYTES_PER_POINT_LAZY if dyc.lazy_velocities else KERNEL_BYTES_PER_POINT_EAGER
    pts = run.nx * run.ny * run.nz
    peak, how = measured_peak_gbs()
    kernels = {}
    for ki, n in enumerate(KERNEL_NAMES):
        ms = float(np.mean([per_stage[st][ki] for st in per_stage]))
        if ms <= 0:
            continue
        bytes_pt = float(np.mean([bpp[min(st, 2)][ki] for st in per_stage]))
        gbs = bytes_pt * pts / (ms * 1e-3) / 1e9
        kernels[n] = {"ms_per_launch": ms, "achieved": gbs, "frac": gbs / peak,
                      "algorithmic_bytes_per_launch": bytes_pt * pts,
                      "algorithmic_bytes_per_point_by_stage": [bpp[min(st, 2)][ki] for st in sorted(per_stage)],
                      "ms_by_stage": [float(per_stage[st][ki]) for st in sorted(per_stage)]}
        traffic, src = ncu_traffic(NCU_KERNEL_KEYS[ki])
        if traffic is not None and (run.nx, run.ny, run.nz) == WORKLOADS["c5"]:
            kernels[n]["traffic"], kernels[n]["traffic_source"] = traffic, src
    # the two once-per-step kernels outside the stages, timed alone on the run's fields (0.56 GB
    # each: far larger than L2): velocity diagnosis of the final state (s, su, sv -> u, v = 40 B/pt)
    # and the diagnostics refresh (s -> p, exn, mtg, h = 40 B/pt)
    other = {}
    st = run.state
    diag = run.diag if hasattr(run, "diag") else run.sub.diag
    pt = run.pt if hasattr(run, "pt") else run.sub.pt
    P, EXN, H, MTG, S = ("air_pressure_on_interface_levels", "exner_function_on_interface_levels",
                         "height_on_interface_levels", "montgomery_potential", "air_isentropic_density")
    calls = {"diagnostics refresh (diag_column_kernel)":
             lambda: diag.get_diagnostic_variables(st[S], pt, st[P], st[EXN], st[MTG], st[H])}
    if dyc.lazy_velocities:
        calls["velocity diagnosis (velocity_xy_kernel)"] = lambda: dyc.diagnose_velocities(st)
    for name, fn in calls.items():
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gbs = 40 * pts / (ms * 1e-3) / 1e9
        other[name] = {"ms_per_launch": ms, "achieved": gbs, "frac": gbs / peak,
                       "algorithmic_bytes_per_launch": 40 * pts, "launches_per_step": 1}
    dom = max(kernels, key=lambda n: kernels[n]["ms_per_launch"])
    stage_ms = float(sum(k["ms_per_launch"] for k in kernels.values()))
    stage_gbs = BYTES_PER_POINT_STAGE * pts / (stage_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved"], "peak": peak,
            "unit": "GB/s", "frac": kernels[dom]["frac"],
            "traffic": kernels[dom].get("traffic"), "traffic_source": kernels[dom].get("traffic_source"),
            "peak_source": how, "ms_per_launch": kernels[dom]["ms_per_launch"],
            "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes_per_launch"],
            "launches_averaged": f"the {nst} RK stages of a step (their compulsory traffic differs: "
                                 "intermediate stages neither read nor write u, v)",
            "kernels": kernels, "other_kernels": other,
            "fused_stage": {"ms": stage_ms, "algorithmic_bytes": BYTES_PER_POINT_STAGE * pts,
                            "achieved": stage_gbs, "frac": stage_gbs / peak,
                            "note": "112 B/pt (SURVEY.md 8d) over the three kernels of one RK stage, mean of the stages"}}


# ------------------------------------------------------------------ the other configurations, briefly
def aux_measurements(args):
    """Short measurements of BASELINE configs[1], [2], [3] after the timed headline region, so
    that one driver run carries every configuration.  Each entry:
    ms/step on the device (CUDA events), launches per step, fraction of the HBM roofline where a
    per-point traffic figure exists.  A failure is reported in place, never raised."""
    import torch

    from tasmania_b200 import lib as tblib

    peak, _ = measured_peak_gbs()
    out = {}

    def timed(step, steps, warmup):
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = tblib.launch_count()
        ev0.record()
        for _ in range(steps):
            step()
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / steps, (tblib.launch_count() - n0) / steps

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as exc:  # noqa: BLE001  (reported, the headline line must still print)
            out[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        torch.cuda.empty_cache()

    def c2():
        nx, ny, nz = WORKLOADS["c2"]
        res = {"workload": workload_name("c2").replace(" per GPU", "")}
        for graph in (False, True):
            run = DryRun(nx, ny, nz, graph=graph)
            ms, launches = timed(run.step, 50, 8)
            for n in run.out_names:
                if not bool(torch.isfinite(run.state[n].t).all()):
                    raise RuntimeError(f"{n} is not finite")
            key = "cuda_graphs" if graph else "eager"
            res[key] = {"ms_per_step": ms, "Mpts_steps_per_s": nx * ny * nz / ms / 1e3,
                        "hbm_frac_step": BYTES_PER_POINT_STEP * nx * ny * nz / (ms * 1e-3) / 1e9 / peak,
                        "launches_per_step": launches if not graph else None}
            del run
        return res

    def c3():
        from tasmania_b200.isentropic_moist import IsentropicMoistSUS, moist_mountain_case

        nx, ny, nz = 256, 256, 60
        grid, np_state = moist_mountain_case(nx, ny, nz, topo_seconds=1800.0, max_height=500.0,
                                             relative_humidity=0.95, seed=True,
                                             half_width_km=(1.1 * (nx - 1), 1.1 * (ny - 1)))
        model = IsentropicMoistSUS(grid, np_state, timedelta(seconds=5))
        ms, launches = timed(model.step, 20, 5)
        for n, v in model.state.items():
            if n != "time" and not bool(torch.isfinite(v.t).all()):
                raise RuntimeError(f"{n} is not finite")
        return {"workload": "moist isentropic + Kessler + sedimentation, sequential-update splitting, "
                            f"{nx}x{ny}x{nz}, fp64", "ms_per_step": ms,
                "Mpts_steps_per_s": nx * ny * nz / ms / 1e3, "launches_per_step": launches}

    def c4():
        from tasmania_b200.diffusion_dwarf import DiffusionDwarfRun

        nx, ny, nz = WORKLOADS["c4"]
        run = DiffusionDwarfRun(nx, ny, nz)
        ms, launches = timed(run.step, 5, 3)
        k_ms, _ = timed(lambda: run.diffusion(run.phi, run.tnd), 5, 2)
        if not bool(torch.isfinite(run.phi.t).all()):
            raise RuntimeError("phi is not finite")
        pts = nx * ny * nz
        return {"workload": f"fourth-order horizontal diffusion dwarf, periodic BC, {nx}x{ny}x{nz}, fp64",
                "ms_per_step": ms, "launches_per_step": launches,
                "step": "diffusion (16 B/pt) + phi update (24 B/pt) + periodic halo",
                "hbm_frac_step": 40 * pts / (ms * 1e-3) / 1e9 / peak,
                "diffusion_kernel": {"ms_per_launch": k_ms, "achieved": 16 * pts / (k_ms * 1e-3) / 1e9,
                                     "frac": 16 * pts / (k_ms * 1e-3) / 1e9 / peak, "unit": "GB/s",
                                     "kernel": os.environ.get("TB200_DIFF_IMPL", "march2")}}

    guarded("c2_dry_161x161x60", c2)
    guarded("c3_moist_256x256x60", c3)
    guarded("c4_diffusion_4096x4096x64", c4)
    return out


def end_to_end(run, args, world, barrier, distributed):
    """Same metric through the public host-buffer API (tasmania_b200.pipeline): every step
    uploads the stage inputs (s, su, sv, u, v, mtg) from pinned host memory, runs the RK step +
    diagnostics refresh and downloads the stepped fields (s, su, sv, u, v) into pinned host
    memory; uploads / computation / downloads of consecutive steps overlap on three streams."""
    import torch

    from tasmania_b200.pipeline import HostStreamedDryCore, flat

    diag = run.diag if hasattr(run, "diag") else run.sub.diag
    pt = run.pt if hasattr(run, "pt") else run.sub.pt
    dt = run.dt if hasattr(run, "dt") else run.sub.dt
    pipe = HostStreamedDryCore(run.dyc, diag, pt, dt, prognostic_only=bool(run.dyc.lazy_velocities))
    host_in = pipe.host_buffers(pipe.names_in)
    for n in pipe.names_in:
        host_in[n].copy_(flat(run.state[n]))
    host_out = pipe.host_buffers(pipe.names_out)
    steps = max(4, min(args.steps, 8))
    pipe.step(host_in, host_out)  # warm-up (allocations of the dycore's stage buffers)
    pipe.join()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        pipe.step(host_in, host_out)
    pipe.join()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if distributed:
        import torch.distributed as dist

        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    nbytes = lambda d: int(sum(t.numel() * 8 for t in d.values())) * world  # noqa: E731  (whole job)
    pts = run.nx * run.ny * run.nz
    return {"value": pts * steps * world / (ms * 1e-3) / 1e6, "unit": "Mpts*steps/s",
            "h2d_bytes_per_step": nbytes(host_in), "d2h_bytes_per_step": nbytes(host_out),
            "steps": steps, "pinned_buffers_numa_node": getattr(pipe, "numa_node", None),
            "api": "tasmania_b200.pipeline.HostStreamedDryCore.step (prognostic fields "
                                   "s, su, sv over PCIe; Montgomery potential and velocities diagnosed on the device)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true",
                    help="skip the short measurements of configs[1..3] after the headline (JSON key 'aux')")
    ap.add_argument("--graph", action="store_true",
                    help="replay the step as CUDA graphs (tasmania_b200.graphs): pays on the "
                         "launch-bound grids (c2), irrelevant at c5 where a step is 8.5 ms of kernels")
    ap.add_argument("--overlap", action="store_true",
                    help="multi-GPU: run the halo exchange on a side stream under the interior "
                         "blocks of the momentum kernel (measured slower at 8 GPUs: 12.19 vs 11.94 "
                         "ms/step, the split launches cost more than the hidden exchange)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload.startswith("c4"):
        run_c4(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
