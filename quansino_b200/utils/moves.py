"""``src/quansino/utils/moves.py:19-102``: move bookkeeping."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any


@dataclass
class MoveStorage:
    """A move, its criteria and its scheduling parameters."""

    move: Any
    criteria: Any
    interval: int
    probability: float
    minimum_count: int

    def to_dict(self) -> dict[str, Any]:
        return {
            "name": self.__class__.__name__,
            "kwargs": {
                "move": self.move.to_dict(),
                "criteria": self.criteria.to_dict(),
                "interval": self.interval,
                "probability": self.probability,
                "minimum_count": self.minimum_count,
            },
        }
