"""``ObserverManager`` (``src/quansino/io/file.py:14-130``): keeps the close callbacks of observers and other file resources, runs
them -- last registered first -- when it is closed, when its ``with`` block ends, or at interpreter exit.  The drivers of this
package keep their observers in ``MonteCarlo.observers`` and close them in ``close()``; the manager is offered to user code that
used it directly."""

from __future__ import annotations

import atexit
from typing import Any, Callable
from warnings import warn


class ObserverManager:
    def __init__(self) -> None:
        self._callbacks: list[Callable[[], Any]] = []
        self.observers: dict[str, Any] = {}
        atexit.register(self.close)

    def __enter__(self):
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    def register(self, resource: Callable[[], Any]) -> Callable[[], Any]:
        """Remember a callback (typically ``file.close``); returns it, so the call can be used as a decorator."""
        self._callbacks.append(resource)
        return resource

    def close(self) -> None:
        """Run the callbacks in reverse order of registration; a failing one does not keep the others from running."""
        while self._callbacks:
            callback = self._callbacks.pop()
            try:
                callback()
            except Exception as err:  # noqa: BLE001  (closing must go on)
                warn(f"closing a registered resource failed: {err}", RuntimeWarning, 2)
        atexit.unregister(self.close)

    def attach_observer(self, name: str, observer) -> None:
        self.observers[name] = observer
        self.register(observer.close)

    def detach_observer(self, name: str, close: bool = False) -> None:
        observer = self.observers.pop(name, None)
        if observer is None:
            warn(f"`Observer` '{name}' not found when deleting.", UserWarning, 2)
        elif close:
            observer.close()


__all__ = ["ObserverManager"]
