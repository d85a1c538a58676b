# -*- coding: utf-8 -*-
"""The horizontal-diffusion dwarf of BASELINE.json configs[3] on the b200 backend, as SURVEY.md
section 8d defines it: a field on a doubly periodic grid, repeatedly

    tnd = HorizontalDiffusion("fourth_order")(phi)     dwarfs/subclasses/horizontal_diffusers/fourth_order.py:L71-L124
    phi = phi + dt * tnd                               the `fma` stencil, stencil_definitions/math.py:L59-L63
    Periodic.enforce_field(phi)                        domain/subclasses/horizontal_boundaries/periodic.py:L98-L114

on the numerical grid ``(nx + 2 nb, ny + 2 nb, nz)`` the reference's ``Periodic`` boundary
builds.  Three launches per application, 16 + 24 B/pt of algorithmic traffic (diffusion: read
phi, write tnd; update: read phi and tnd, write phi) plus the halo frame.
"""
from __future__ import annotations

from tasmania_b200 import stencils, storage
from tasmania_b200.boundary import Periodic
from tasmania_b200.dwarfs import HorizontalDiffusion


class DiffusionDwarfRun:
    def __init__(self, nx, ny, nz, phi=None, *, diffusion_type="fourth_order", nb=2, dx=1.0, dy=1.0,
                 diffusion_coeff=0.5, diffusion_coeff_max=1.0, diffusion_damp_depth=15, dt=0.05,
                 seed=20261018, device=None):
        """``phi``: the physical field ``(nx, ny, nz)`` (host or device array); None draws a
        standard-normal field on the device (synthetic benchmark input)."""
        self.nx, self.ny, self.nz, self.nb, self.dt = nx, ny, nz, nb, dt
        self.hb = Periodic(nx, ny, nz, nb)
        if phi is None:
            import torch

            phys = storage.zeros((nx, ny, nz), device=device)
            gen = torch.Generator(device=phys.t.device)
            gen.manual_seed(seed)
            phys.t.normal_(generator=gen)
        else:
            phys = storage.as_storage(phi, device=device)
        self.phi = self.hb.get_numerical_field(phys)
        del phys
        self.shape = tuple(self.phi.shape)
        self.diffusion = HorizontalDiffusion.factory(
            diffusion_type, self.shape, dx, dy, diffusion_coeff, diffusion_coeff_max,
            diffusion_damp_depth, nb)
        self.tnd = storage.zeros(self.shape, device=device)
        self.nstep = 0

    def step(self):
        self.diffusion(self.phi, self.tnd, overwrite_output=True)
        stencils.fma_fields([self.phi], [self.phi], [self.tnd], self.dt, origin=(0, 0, 0),
                            domain=self.shape)
        self.hb.enforce_field(self.phi)
        self.nstep += 1
        return self.phi

    def physical_field(self):
        return self.hb.get_physical_field(self.phi)
