"""Displacement operations (``src/quansino/operations/displacement.py``).

Draw order inside the kernel (``csrc/sweep_kernel.cuh``), identical to the reference:
Ball r, phi, cos(theta) (``:146-149``); Sphere phi, cos(theta) (``:103-104``); Box x, y, z (``:67-68``);
Translation f0, f1, f2 (``:173``)."""

from __future__ import annotations

from typing import Any

from quansino_b200 import _lib
from quansino_b200.operations.core import BaseOperation


class DisplacementOperation(BaseOperation):
    def __init__(self, step_size: float = 1.0) -> None:
        self.step_size = step_size

    def to_dict(self) -> dict[str, Any]:
        return {**super().to_dict(), "kwargs": {"step_size": self.step_size}}


class Box(DisplacementOperation):
    """Uniform in [-step_size, step_size]^3."""

    op_code = _lib.OP_BOX


class Sphere(DisplacementOperation):
    """On the sphere of radius step_size."""

    op_code = _lib.OP_SPHERE


class Ball(DisplacementOperation):
    """Inside the ball of radius step_size; the radius is uniform in [0, step_size] (not volume-uniform), as in the
    reference."""

    op_code = _lib.OP_BALL


class Translation(BaseOperation):
    """Move the selected atoms to a uniform point of the cell (``operations/displacement.py:156-175``): exchange insertions, and
    displacement moves of label groups (the centroid lands on the point)."""

    op_code = _lib.OP_TRANSLATION


class Rotation(BaseOperation):
    """Rigid rotation of the selected label group about its centre of mass (``operations/displacement.py:178-200``): three
    angles ``uniform(0, 2 pi)`` handed to ASE's ``euler_rotate`` -- which reads them as DEGREES, so the rotation is at most
    6.3 degrees per axis; the quirk is reproduced."""

    op_code = _lib.OP_ROTATION


class TranslationRotation(BaseOperation):
    """``Translation`` + ``Rotation`` of the selected label group (``operations/displacement.py:203-227``): the centroid goes to a
    uniform point of the cell, the group turns about its centre of mass; the translation draws first.  Displacement moves of label
    groups only (molecular insertion is not lowered yet)."""

    op_code = _lib.OP_TRANSLATION_ROTATION
