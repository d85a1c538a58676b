# -*- coding: utf-8 -*-
"""The kernels' own correctly rounded division (csrc/common.cuh: qdiv, used by the velocity
diagnoses inside the fused stage kernels) against the compiler's IEEE division: bit for bit on
2^28 generated operand pairs -- the data range, pairs that cross the fast path's guard on either
side, divisor mantissas of all ones, quotients next to 1, powers of two, zero / subnormal
numerators (tb200_selftest_division)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 20261018])
def test_qdiv_equals_ieee_division_bitwise(seed):
    import torch

    from tasmania_b200 import lib

    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    n = 1 << 28
    lib.check(lib.load().tb200_selftest_division(n, seed, bad.data_ptr(), lib.current_stream()),
              "tb200_selftest_division")
    torch.cuda.synchronize()
    assert int(bad.item()) == 0, f"{int(bad.item())} of {n} quotients differ from a / b"
