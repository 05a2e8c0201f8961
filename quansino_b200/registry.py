"""Class registry for (de)serialisation (``src/quansino/registry.py``): ``to_dict()`` stores ``__class__.__name__``,
``from_dict`` looks the class up again.  The mirrored classes of this package are registered on import; user classes can be
registered the same way (they still cannot run on the device)."""

from __future__ import annotations

from typing import Callable

_class_registry: dict[str, type] = {}


def register_class(cls: type, class_name: str | None = None) -> type:
    _class_registry[class_name or cls.__name__] = cls
    return cls


def register(class_name: str | None = None) -> Callable[[type], type]:
    def decorator(cls: type) -> type:
        return register_class(cls, class_name)

    return decorator


def get_class(class_name: str) -> type:
    _ensure_builtin()
    if class_name not in _class_registry:
        raise KeyError(f"Class `{class_name}` not registered in the global registry. Available classes: {sorted(_class_registry)}")
    return _class_registry[class_name]


def get_typed_class(class_name: str, expected_base: type | None = None) -> type:
    cls = get_class(class_name)
    if expected_base is not None and isinstance(expected_base, type) and not issubclass(cls, expected_base):
        raise TypeError(f"Class `{class_name}` is not a {expected_base.__name__}")
    return cls


def get_class_name(cls: type) -> str | None:
    """The registered name of a class, or None (``registry.py:105-124``)."""
    _ensure_builtin()
    for name, registered in _class_registry.items():
        if registered is cls:
            return name
    return None


def get_class_registry() -> dict[str, type]:
    _ensure_builtin()
    return dict(_class_registry)


def _ensure_builtin() -> None:
    if "Ball" in _class_registry:
        return
    from quansino_b200 import integrators, mc, moves, operations
    from quansino_b200.utils.moves import MoveStorage

    for module in (operations, moves, mc, integrators):
        for name in getattr(module, "__all__", []):
            obj = getattr(module, name)
            if isinstance(obj, type):
                register_class(obj)
    register_class(MoveStorage)
