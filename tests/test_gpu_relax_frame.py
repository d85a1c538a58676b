# -*- coding: utf-8 -*-
"""Relaxed.enforce_raw through ``tb200_relax_frame`` (all fields in one launch over the frame where
gamma != 0) against the reference-shaped path (one full-box ``irelax`` per field) and against the
oracle: bit-exact, on a whole domain and on sub-domain windows of a decomposed grid."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NAMES = ("air_isentropic_density", "x_momentum_isentropic", "y_momentum_isentropic",
         "x_velocity_at_u_locations", "y_velocity_at_v_locations",
         "mass_fraction_of_water_vapor_in_air", "mass_fraction_of_cloud_liquid_water_in_air",
         "mass_fraction_of_precipitation_water_in_air", "height_on_interface_levels",
         "air_pressure_on_interface_levels")


def _run(hb, fields, mode):
    import tasmania_b200 as tb

    state = {n: tb.as_storage(v) for n, v in fields.items()}
    old = os.environ.get("TB200_RELAX")
    os.environ["TB200_RELAX"] = mode
    try:
        n0 = tb.lib.launch_count()
        hb.enforce_raw(state, {n: {} for n in NAMES[:-1]})  # the last field is not to be touched
        launches = tb.lib.launch_count() - n0
    finally:
        if old is None:
            del os.environ["TB200_RELAX"]
        else:
            os.environ["TB200_RELAX"] = old
    return {n: tb.to_numpy(v) for n, v in state.items()}, launches


@pytest.mark.parametrize("case", ["whole_nr6", "whole_nr8", "west_edge", "corner", "interior"])
def test_frame_relaxation_is_bitwise_the_full_one(case):
    from oracle import boundary as ob
    from tasmania_b200.boundary import Relaxed

    rng = np.random.default_rng(7)
    nx, ny, nz, nb = 45, 38, 7, 3
    kw = {"whole_nr6": dict(nr=6), "whole_nr8": dict(nr=8),
          "west_edge": dict(nr=6, global_extent=(400, 300), offset=(0, 120)),
          "corner": dict(nr=6, global_extent=(400, 300), offset=(400 - nx, 300 - ny)),
          "interior": dict(nr=6, global_extent=(400, 300), offset=(150, 120))}[case]
    hb = Relaxed(nx, ny, nz, nb, **kw)
    shape = (nx + 1, ny + 1, nz + 1)
    fields = {n: rng.standard_normal(shape) for n in NAMES}
    hb.reference_state = {n: rng.standard_normal(shape) for n in NAMES}
    full, n_full = _run(hb, fields, "full")
    frame, n_frame = _run(hb, fields, "frame")
    assert n_full == 9 and n_frame == (0 if case == "interior" else 2)  # 9 fields: 8 + 1
    for n in NAMES:
        np.testing.assert_array_equal(frame[n], full[n], err_msg=n)
    np.testing.assert_array_equal(frame[NAMES[-1]], fields[NAMES[-1]])
    if case.startswith("whole"):
        ohb = ob.Relaxed(nx, ny, nz, nb, kw["nr"])
        import tasmania_b200 as tb

        ohb.reference_state = {n: tb.to_numpy(v) for n, v in hb.reference_state.items()}
        want = {n: v.copy() for n, v in fields.items()}
        ohb.enforce_raw(want, NAMES[:-1])
        for n in NAMES:
            np.testing.assert_array_equal(frame[n], want[n], err_msg=n)
        assert not np.array_equal(frame[NAMES[0]], fields[NAMES[0]])
