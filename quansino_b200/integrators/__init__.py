"""Mirror of ``quansino.integrators`` for the GPU path."""

from quansino_b200.integrators.displacement import BaseIntegrator, Verlet

__all__ = ["BaseIntegrator", "Verlet"]
