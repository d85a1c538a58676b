# -*- coding: utf-8 -*-
"""Host-side regression check for the GPU-less container: run the bodies of GPU parity tests with the
oracle-backed C-ABI stub in place of the library (tests/abi_oracle.py).  Whatever a test compares
against fixtures / the oracle then has to come out bit-exact or within its own tolerance; a Python-level
regression in the mirrors or the marshalling shows up here without a GPU.

    python experiments/gpu_tests_under_stub.py
"""
import itertools
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests.abi_oracle import OracleStub  # noqa: E402
from tests.abi_stub import stubbed_library  # noqa: E402


def main():
    from tests import test_gpu_isentropic as tg

    dry = ("isen_dry_rk3_5th", "isen_dry_rk3_3rd", "isen_dry_rk3_cen", "isen_dry_fe_upw",
           "isen_dry_rk3_5th_periodic", "isen_dry_rk3_3rd_periodic", "isen_dry_fe_cen_periodic")
    moist = ("isen_moist_rk3_5th", "isen_moist_fe_3rd", "isen_moist_rk3_5th_periodic")
    modes = ((False, False), (True, False), (True, True))
    jobs = [(tg.test_dry_dycore_vs_reference_fixture, (c, f, l)) for c, (f, l) in itertools.product(dry, modes)]
    jobs += [(tg.test_moist_dycore_vs_reference_fixture, (c, f, l)) for c, (f, l) in itertools.product(moist, modes)]
    jobs += [(tg.test_fused_moist_stage_equals_stencil_path_bitwise, ()),
             (tg.test_fused_equals_stencil_path_bitwise, ())]
    failed = 0
    for fn, args in jobs:
        try:
            with stubbed_library(OracleStub):
                fn(*args)
        except Exception:  # noqa: BLE001
            failed += 1
            print("FAILED", fn.__name__, args)
            traceback.print_exc()
    print(f"GPU-TESTS-UNDER-STUB {len(jobs) - failed}/{len(jobs)} ok")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
