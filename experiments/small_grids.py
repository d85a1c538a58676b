#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""BASELINE configs[1] (dry, 161x161x60) and configs[2] (moist SUS model, 256x256x60) on one GPU:
device time per step (CUDA events), host enqueue time per step (time.perf_counter around the same
loop before the synchronize) and kernel launches per step.  At these sizes every field fits the
126 MB L2 several times over and a step is a chain of 10 / ~130 short kernels, so the question is
whether the device or the Python/ctypes launch path is the limiter.

    python experiments/small_grids.py [--steps 50]
"""
import argparse
import json
import os
import sys
import time
from datetime import timedelta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(step, steps, warmup=5):
    import torch

    from tasmania_b200 import lib

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.launch_count()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return {"device_ms_per_step": ev0.elapsed_time(ev1) / steps, "host_enqueue_ms_per_step": host / steps * 1e3,
            "launches_per_step": (lib.launch_count() - n0) / steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--only", default=None, choices=("c2", "c3"), help="for ncu launch lists")
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    import torch

    import bench
    from tasmania_b200.isentropic_moist import IsentropicMoistSUS
    from tests import helpers as hp

    out = {}
    if args.only is not None:
        if args.only == "c2":
            nx, ny, nz = bench.WORKLOADS["c2"]
            step = bench.DryRun(nx, ny, nz).step
        else:
            nx, ny, nz = 256, 256, 60
            grid, np_state = hp.moist_case(nx, ny, nz, topo_seconds=1800.0, max_height=500.0,
                                           relative_humidity=0.95, seed=True,
                                           half_width_km=(1.1 * (nx - 1), 1.1 * (ny - 1)))
            step = IsentropicMoistSUS(grid, np_state, timedelta(seconds=5)).step
        print(json.dumps({args.only: timed(step, args.steps, warmup=args.warmup)}))
        return
    nx, ny, nz = bench.WORKLOADS["c2"]
    run = bench.DryRun(nx, ny, nz)
    r = timed(run.step, args.steps)
    r["Mpts_steps_per_s"] = nx * ny * nz / r["device_ms_per_step"] / 1e3
    r["hbm_frac_step"] = (bench.BYTES_PER_POINT_STEP * nx * ny * nz / (r["device_ms_per_step"] * 1e-3) / 1e9
                          / bench.measured_peak_gbs()[0])
    out["c2_dry_161x161x60"] = r
    del run
    run = bench.DryRun(nx, ny, nz, graph=True)
    r = timed(run.step, args.steps, warmup=6)
    r["Mpts_steps_per_s"] = nx * ny * nz / r["device_ms_per_step"] / 1e3
    r["hbm_frac_step"] = (bench.BYTES_PER_POINT_STEP * nx * ny * nz / (r["device_ms_per_step"] * 1e-3) / 1e9
                          / bench.measured_peak_gbs()[0])
    r["graphs"] = run.loop.period
    out["c2_dry_161x161x60_cuda_graphs"] = r
    del run
    nx, ny, nz = 256, 256, 60
    grid, np_state = hp.moist_case(nx, ny, nz, topo_seconds=1800.0, max_height=500.0,
                                   relative_humidity=0.95, seed=True,
                                   half_width_km=(1.1 * (nx - 1), 1.1 * (ny - 1)))
    model = IsentropicMoistSUS(grid, np_state, timedelta(seconds=5))
    r = timed(model.step, args.steps)
    r["Mpts_steps_per_s"] = nx * ny * nz / r["device_ms_per_step"] / 1e3
    for n, v in model.state.items():
        if n != "time" and not bool(torch.isfinite(v.t).all()):
            raise RuntimeError(f"{n} is not finite")
    out["c3_moist_sus_256x256x60"] = r
    from tasmania_b200.graphs import GraphedLoop

    loop = GraphedLoop(model, eager_steps=0)
    r = timed(loop.step, args.steps, warmup=14)
    r["Mpts_steps_per_s"] = nx * ny * nz / r["device_ms_per_step"] / 1e3
    r["graphs"] = loop.period
    for n, v in model.state.items():
        if n != "time" and not bool(torch.isfinite(v.t).all()):
            raise RuntimeError(f"{n} is not finite")
    out["c3_moist_sus_256x256x60_cuda_graphs"] = r
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
