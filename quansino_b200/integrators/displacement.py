"""Integrators (``src/quansino/integrators/displacement.py:16-81``)."""

from __future__ import annotations

from typing import Any

from quansino_b200 import _lib

fs = 0.09822694788464063  # ase.units.fs (CODATA 2014): 1e-15 * 1e10 * sqrt(_e / _amu)


class BaseIntegrator:
    op_code = -1

    def integrate(self, context) -> None:
        raise NotImplementedError(
            f"{type(self).__name__}.integrate runs inside the CUDA trajectory kernel; it cannot be evaluated on the host"
        )

    def to_dict(self) -> dict[str, Any]:
        return {"name": self.__class__.__name__}

    @classmethod
    def from_dict(cls, data: dict[str, Any]):
        return cls(**data.get("kwargs", {}))


class Verlet(BaseIntegrator):
    """Velocity Verlet, ``dt`` in femtoseconds, ``max_steps`` force evaluations per trajectory."""

    op_code = _lib.OP_VERLET

    def __init__(self, dt: float = 1.0, max_steps: int = 100, apply_constraints: bool = True) -> None:
        self.dt = dt * fs
        self.max_steps = max_steps
        self.apply_constraints = apply_constraints

    def lower(self):
        if type(self).integrate is not BaseIntegrator.integrate:
            raise NotImplementedError("user-defined integrators cannot execute on the device")
        if not self.apply_constraints:
            raise NotImplementedError(
                "Verlet(apply_constraints=False) is not available on the GPU path (the kernel reproduces the "
                "half-step momentum re-derivation of integrators/displacement.py:74-75)"
            )
        return self.op_code, float(self.dt), int(self.max_steps)

    def to_dict(self) -> dict[str, Any]:
        return {**super().to_dict(), "kwargs": {"dt": self.dt / fs, "max_steps": self.max_steps, "apply_constraints": self.apply_constraints}}
