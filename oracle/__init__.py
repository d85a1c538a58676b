# -*- coding: utf-8 -*-
"""CPU oracle for the tasmania stencil hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A numpy restatement of the reference's *numpy backend* for the path named by
BASELINE.json:north_star, each function citing the
reference file:line it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package;
``tasmania_b200`` (the product) never does and fails loudly when its CUDA library is
missing.

Parity pin: the reference ships no golden vectors, and it cannot be
imported as shipped (gt4py / sympl fork / pint / xarray are absent).  The oracle is
therefore pinned against outputs of *the reference's own numpy code executed in place*
from ``/root/reference`` (``tests/golden/refload.py`` + ``tests/golden/generate_golden.py``,
both committed), stored as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks
every oracle function against those fixtures bit-for-bit (no tolerance: same numpy, same
operation order), so "parity pinned: yes, by execution of the reference".

Operation order is kept exactly as in the reference (``u / 60.0 * (...)``, divide by
``dx`` rather than multiply by a reciprocal, sequential vertical scans) because the CUDA
kernels are compared to this oracle at 1e-12 relative.
"""

from oracle import (boundary, burgers, dwarfs, fluxes, isentropic, isentropic_physics,  # noqa: F401
                    microphysics)
