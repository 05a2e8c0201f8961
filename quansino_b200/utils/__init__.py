"""Mirror of the hot-path parts of ``quansino.utils``."""

from quansino_b200.utils.dynamics import maxwell_boltzmann_distribution
from quansino_b200.utils.moves import MoveStorage

__all__ = ["MoveStorage", "maxwell_boltzmann_distribution"]
