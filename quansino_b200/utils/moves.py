"""``src/quansino/utils/moves.py:19-102``: move bookkeeping."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any


@dataclass
class MoveStorage:
    """A move, its criteria and its scheduling parameters."""

    def __class_getitem__(cls, item):
        """``Isobaric[Move, Criteria](...)``, ``CellMove[OperationType, ContextType]``: the reference's classes are ``Generic``; the
        parameters carry no run-time meaning here either."""
        return cls


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

    @classmethod
    def from_dict(cls, data: dict[str, Any]):
        """``utils/moves.py:62-102``: move and criteria are rebuilt through the registry."""
        from quansino_b200.registry import get_class

        kw = dict(data.get("kwargs", {}))
        kw["move"] = get_class(kw["move"]["name"]).from_dict(kw["move"])
        kw["criteria"] = get_class(kw["criteria"]["name"]).from_dict(kw["criteria"])
        return cls(**kw)
