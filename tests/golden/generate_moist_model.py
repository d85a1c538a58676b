# -*- coding: utf-8 -*-
"""Golden fixture of BASELINE configs[2] produced by the REFERENCE ITSELF, run in place from
/root/reference (numpy backend): five steps of the moist benchmark loop -- the reference's moist
dynamical core stage and its physics suite, chained by its own couplers and steppers exactly as
tests/test_moist_physics_reference.py does -- on the seeded 21 x 19 x 10 mountain-flow case.

    python tests/golden/generate_moist_model.py        -> tests/golden/moist_model.npz

The fixture travels to the GPU box, where /root/reference does not exist
(tests/test_gpu_moist_model.py::test_moist_model_vs_reference_fixture).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import test_moist_physics_reference as ref  # noqa: E402

NSTEPS = 5


def main():
    model, ost, domain, dt = ref._setup(0)
    physics, dycore = ref.reference_physics(model, domain), ref.reference_dycore(model, domain)
    grid = domain.numerical_grid
    state = {n: ref.da(v.copy(), n) for n, v in ost.items() if n != "time"}
    state["time"] = ost["time"]
    out = {"init_" + n: v.data.copy() for n, v in state.items() if n != "time"}
    with np.errstate(divide="ignore", invalid="ignore"):
        for step in range(NSTEPS):
            grid.update_topography((step + 1) * dt)
            for n, v in dycore(ref.raw(state), dt).items():
                state[n] = ref.da(v, n)
            physics(state, dt)
    out.update({"final_" + n: v.data.copy() for n, v in state.items() if n != "time"})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "moist_model.npz")
    np.savez_compressed(
        path, dims=np.array([ref.NX, ref.NY, ref.NZ, ref.NB, 6, NSTEPS, 4]),
        params=np.array([dt.total_seconds(), 500.0, 60.0, 0.98]),  # dt, mountain height, growth time, RH
        **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
