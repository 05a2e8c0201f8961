"""Operation base classes (``src/quansino/operations/core.py:16-139``).

On the GPU path an Operation is a *description* that lowers to ``(qmc_op, parameter)``; the draw itself happens
in the sweep kernel from the shared Philox stream, in the draw order of the reference."""

from __future__ import annotations

from typing import Any


class BaseOperation:
    """Mirror of ``quansino.operations.core.BaseOperation``."""

    op_code: int = -1

    def calculate(self, context) -> Any:
        raise NotImplementedError(
            f"{type(self).__name__}.calculate runs inside the CUDA sweep kernel; it cannot be evaluated on the host "
            "(quansino_b200 has no CPU fallback)"
        )

    def lower(self) -> tuple[int, float]:
        """``(qmc_op code, step / max_value)``; custom subclasses cannot run on the device."""
        if type(self).calculate is not BaseOperation.calculate or self.op_code < 0:
            raise NotImplementedError(
                f"operation {type(self).__name__} is not available on the GPU path (user-defined Operation subclasses "
                "cannot execute on the device)"
            )
        return self.op_code, float(getattr(self, "step_size", getattr(self, "max_value", 0.0)))

    def to_dict(self) -> dict[str, Any]:
        return {"name": self.__class__.__name__}

    @classmethod
    def from_dict(cls, data: dict[str, Any]):
        return cls(**data.get("kwargs", {}))

    def __add__(self, other):
        """``operations/core.py:72-93``."""
        from quansino_b200.operations.composite import CompositeOperation

        if isinstance(other, CompositeOperation):
            return CompositeOperation([self, *other.operations])
        return CompositeOperation([self, other])

    def __mul__(self, n: int):
        """``operations/core.py:95-116``."""
        from quansino_b200.operations.composite import CompositeOperation

        if not isinstance(n, int) or isinstance(n, bool) or n < 1:
            raise ValueError("The number of times the move is repeated must be a positive, non-zero integer.")
        return CompositeOperation([self] * n)

    __rmul__ = __mul__
