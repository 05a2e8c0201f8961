"""``src/quansino/integrators/core.py``: the integrator base class (it lives next to ``Verlet`` in this package)."""

from quansino_b200.integrators.displacement import BaseIntegrator

__all__ = ["BaseIntegrator"]
