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
        raise NotImplementedError("CompositeOperation is not available on the GPU path (SURVEY.md section 8f)")
