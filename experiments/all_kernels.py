#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Every kernel family of the path on fields LARGER than L2, for the per-kernel HBM evidence
north_star asks for (`ncu --set full` over this script; profiles/README.md, round 2):

    python experiments/all_kernels.py moist   [--size 768 768 64] [--steps 2]
        the moist SUS model of configs[2]: fused moist stage (s-step, tracer kernel, scans, momentum),
        velocity pass, diagnostics, Coriolis, smoothing, Smagorinsky, Kessler, saturation adjustment,
        vertical advection with fused stage update, fall velocity, sedimentation, precipitation,
        fma_fields
    python experiments/all_kernels.py burgers [--size 8192 8192 1]
        Burgers forward-Euler stages (third order) on a Dirichlet box
    python experiments/all_kernels.py halo    [--size 1024 1024 64]
        the peer-store halo exchange between two sub-domains living on this one GPU
Without ncu it prints device ms per step (CUDA events)."""
import argparse
import json
import os
import sys
from datetime import timedelta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(step, steps, warmup):
    import torch

    from tasmania_b200 import lib

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.launch_count()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    return {"device_ms_per_step": ev0.elapsed_time(ev1) / steps, "launches_per_step": (lib.launch_count() - n0) / steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=("moist", "burgers", "halo"))
    ap.add_argument("--size", type=int, nargs=3, default=None)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    import numpy as np
    import torch

    import tasmania_b200 as tb

    if args.what == "moist":
        from tasmania_b200.isentropic_moist import IsentropicMoistSUS, moist_mountain_case

        nx, ny, nz = args.size or (768, 768, 64)
        grid, np_state = moist_mountain_case(nx, ny, nz, topo_seconds=1800.0, max_height=500.0,
                                             relative_humidity=0.95, seed=True,
                                             half_width_km=(1.1 * (nx - 1), 1.1 * (ny - 1)))
        model = IsentropicMoistSUS(grid, np_state, timedelta(seconds=5))
        r = timed(model.step, args.steps, args.warmup)
        for n, v in model.state.items():
            if n != "time" and not bool(torch.isfinite(v.t).all()):
                raise RuntimeError(f"{n} is not finite")
    elif args.what == "burgers":
        from datetime import datetime

        from tasmania_b200.boundary import Dirichlet
        from tasmania_b200.burgers import BurgersDynamicalCore, ZhaoSolutionFactory
        from tasmania_b200.grid import Grid

        nx, ny, nz = args.size or (8192, 8192, 1)
        assert nz == 1
        grid = Grid((0.0, 1.0), nx, (0.0, 1.0), ny, (0.0, 1.0), 1)
        t0 = datetime(2000, 1, 1)
        zsf = ZhaoSolutionFactory(t0, 0.01)
        state = {n: tb.as_storage(np.asarray(zsf(t0, grid, field_name=n), dtype=np.float64).reshape(nx, ny, 1))
                 for n in ("x_velocity", "y_velocity")}
        state["time"] = t0
        hb = Dirichlet(nx, ny, 1, 2, core=zsf, grid=grid)
        hb.reference_state = state
        dyc = BurgersDynamicalCore(grid, hb, "rk3ws", "third_order")
        box = {"state": state}

        def step():
            out = dyc(box["state"], {}, timedelta(seconds=1e-5))
            box["state"] = {"x_velocity": out["x_velocity"], "y_velocity": out["y_velocity"], "time": out["time"]}

        r = timed(step, args.steps, args.warmup)
    else:
        from tasmania_b200.distributed import InProcessDecomposedRun

        nx, ny, nz = args.size or (1024, 1024, 64)
        run = InProcessDecomposedRun(nx, ny, nz, 2, 1, transport="p2p",
                                     domain_x=(-1.1 * (nx - 1), 1.1 * (nx - 1)), domain_y=(-1.1 * (ny - 1), 1.1 * (ny - 1)))
        r = timed(run.step, args.steps, args.warmup)
        for s in run.subs:
            s.halo.check()
    r["points"] = nx * ny * nz
    r["Mpts_steps_per_s"] = nx * ny * nz / r["device_ms_per_step"] / 1e3
    print(json.dumps({args.what: r}))


if __name__ == "__main__":
    main()
