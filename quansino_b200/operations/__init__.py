"""Mirror of ``quansino.operations`` for the GPU path."""

from quansino_b200.operations.cell import AnisotropicDeformation, DeformationOperation, IsotropicDeformation, ShapeDeformation
from quansino_b200.operations.composite import CompositeOperation
from quansino_b200.operations.core import BaseOperation
from quansino_b200.operations.displacement import Ball, Box, DisplacementOperation, Rotation, Sphere, Translation, TranslationRotation

__all__ = [
    "AnisotropicDeformation", "Ball", "BaseOperation", "Box", "CompositeOperation", "DeformationOperation", "DisplacementOperation",
    "IsotropicDeformation", "Rotation", "ShapeDeformation", "Sphere", "Translation", "TranslationRotation",
]  # fmt: skip
