"""``ObserverManager`` (``src/quansino/io/file.py:14-130``): owns the close callbacks of observers and other file resources; usable
as a context manager, closes what is left at interpreter exit.  The drivers of this package keep their observers in
``MonteCarlo.observers`` and close them in ``close()``; the manager is offered for user code that used it directly."""

from __future__ import annotations

import atexit
from contextlib import ExitStack, suppress
from typing import Any, Callable
from warnings import warn


class ObserverManager:
    def __init__(self) -> None:
        self.exitstack = ExitStack()
        self.observers: dict[str, Any] = {}
        atexit.register(self.close)

    def __enter__(self):
        return self

    def __exit__(self, *args, **kwargs) -> None:
        self.close()

    def close(self) -> None:
        """Run every registered callback (last registered first) and forget the exit hook."""
        with suppress(Exception):
            self.exitstack.close()
            atexit.unregister(self.close)

    def register(self, resource: Callable) -> Any:
        """Register a callback (typically ``file.close``); returns it, like ``ExitStack.callback``."""
        return self.exitstack.callback(resource)

    def attach_observer(self, name: str, observer) -> None:
        self.register(observer.close)
        self.observers[name] = observer

    def detach_observer(self, name: str, close: bool = False) -> None:
        if observer := self.observers.pop(name, None):
            if close:
                observer.close()
        else:
            warn(f"`Observer` '{name}' not found when deleting.", UserWarning, 2)


__all__ = ["ObserverManager"]
