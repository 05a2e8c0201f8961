"""Observer base classes (``src/quansino/io/core.py:12-259``): same constructor arguments, file handling and ``to_dict``."""

from __future__ import annotations

import atexit
import sys
from contextlib import suppress
from io import IOBase
from pathlib import Path
from typing import IO, Any


class Observer:
    """Called by the driver every ``interval`` steps (``interval < 0``: once, at step ``abs(interval)``)."""

    def __init__(self, interval: int = 1) -> None:
        self.interval = interval

    def __call__(self, *args: Any, **kwargs: Any) -> None:
        raise NotImplementedError

    def attach_simulation(self, *args: Any, **kwargs: Any) -> None:
        """Nothing to do by default."""

    def close(self) -> None:
        """Nothing to release by default."""

    def to_dict(self) -> dict[str, Any]:
        return {"name": self.__class__.__name__, "kwargs": {"interval": self.interval}}


class TextObserver(Observer):
    """Writes to a path (opened with ``mode`` / ``encoding``) or to an open stream (``io/core.py:86-259``)."""

    accept_stream: bool = True

    def __init__(self, file: IO | Path | str, interval: int = 1, mode: str = "a", encoding: str | None = None) -> None:
        super().__init__(interval)
        self.mode = mode
        self.encoding = encoding or ("utf-8" if "b" not in mode else None)
        self.file = file

    def __str__(self) -> str:
        if self._file in (sys.stdout, sys.stderr, sys.stdin):
            return f"Stream:{self._file.name}"
        name = getattr(self._file, "name", None)
        if isinstance(name, str) and name not in ("", "."):
            return f"Path:{name}"
        return f"Class:{self._file.__class__.__name__}"

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.__str__()}, mode={self.mode}, encoding={self.encoding}, interval={self.interval})"

    @property
    def file(self) -> IO:
        return self._file

    @file.setter
    def file(self, value: IO | str | Path) -> None:
        if hasattr(self, "_file"):
            self.close()
        if isinstance(value, str):
            value = Path(value)
        if isinstance(value, Path):
            self._file = value.open(mode=self.mode, encoding=self.encoding)
        elif hasattr(value, "read") or hasattr(value, "write") or isinstance(value, IOBase):
            if getattr(value, "closed", False):
                raise ValueError(f"Impossible to link a closed file for '{self.__class__.__name__}'.")
            if not self.accept_stream:
                seekable = value.seekable() if hasattr(value, "seekable") else hasattr(value, "seek")
                if not seekable:
                    raise ValueError(
                        f"{self.__class__.__name__} does not accept non-file streams (non-seekable). Please use a different file type."
                    )
            self._file = value
        else:
            raise TypeError(f"Invalid file type: {type(value)}. Expected str, Path, or file-like object.")
        atexit.register(self.close)

    def close(self) -> None:
        if not hasattr(self, "_file"):
            return
        if self._file not in (sys.stdout, sys.stderr, sys.stdin):
            with suppress(OSError, AttributeError, ValueError):
                self._file.close()
                atexit.unregister(self.close)

    def to_dict(self) -> dict[str, Any]:
        d = super().to_dict()
        d["kwargs"]["mode"] = self.mode
        d["kwargs"]["encoding"] = self.encoding
        return d
