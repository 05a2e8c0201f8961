"""``Logger`` (``src/quansino/io/logger.py:27-335``): one formatted line per call, fields are callables.

Same field API as the reference (``add_field``, ``add_mc_fields``, ``remove_fields``, ``create_header``,
``write_header``).  A batched simulation has one energy per replica, so the convenience fields log batch statistics:
``Epot[eV]`` is the MEAN potential energy over the replicas, and ``add_batch_fields`` adds min / max / mean atom count."""

from __future__ import annotations

import re
from pathlib import Path
from typing import IO, Any, Callable

import numpy as np

from quansino_b200.io.core import TextObserver


def get_auto_header_format(data_format: str) -> str:
    """``utils/strings.py:6-31``: '{:10.3f}' -> '{:>10s}' so that header and data columns line up."""
    return re.sub(r":([<>^])?(\d+)?[^}]*", lambda m: f':{m.group(1) or ">"}{m.group(2) or "10"}s', data_format)


class Logger(TextObserver):
    def __init__(self, logfile: IO | str | Path, interval: int, mode: str = "a", **observer_kwargs: Any) -> None:
        super().__init__(file=logfile, interval=interval, mode=mode, **observer_kwargs)
        self.fields: dict[str | tuple[str, ...], dict[str, Any]] = {}
        self._header_written = False

    def __call__(self) -> None:
        if not self._header_written:  # the drivers add fields after construction; the header goes out with the first line
            self.write_header()
        parts = []
        for key in self.fields:
            f = self.fields[key]
            value = f["function"]()
            parts.append(f["str_format"].format(*value) if f["is_array"] else f["str_format"].format(value))
        self._file.write(" ".join(parts) + "\n")
        self._file.flush()

    def create_header(self) -> str:
        out = []
        for name, f in self.fields.items():
            out.append(f["header_format"].format(*name) if f["is_array"] and isinstance(name, tuple) else f["header_format"].format(name))
        return " ".join(out)

    def write_header(self) -> None:
        self._file.write(f"{self.create_header()}\n")
        self._header_written = True

    def add_field(self, name, function: Callable[[], Any], str_format: str = "{:10.3f}", header_format: str | None = None,
                  is_array: bool = False) -> None:  # fmt: skip
        if isinstance(name, list):
            name = tuple(name)
        self.fields[name] = {
            "function": function,
            "str_format": str_format,
            "header_format": header_format if header_format is not None else get_auto_header_format(str_format),
            "is_array": is_array,
        }

    def add_mc_fields(self, simulation) -> None:
        """Class, Step, Epot[eV] (``io/logger.py:186-212``); Epot is the mean over the replicas of the batch."""
        self.add_field("Class", lambda: simulation.__class__.__name__, "{:<24s}")
        self.add_field("Step", lambda: simulation.step_count, "{:>12d}")
        self.add_field("Epot[eV]", lambda: float(np.mean(simulation.context.last_potential_energy)), "{:>12.4f}")

    def add_batch_fields(self, simulation) -> None:
        """Batch statistics that have no single-system counterpart: energy range and mean atom count."""
        self.add_field("Emin[eV]", lambda: float(np.min(simulation.context.last_potential_energy)), "{:>12.4f}")
        self.add_field("Emax[eV]", lambda: float(np.max(simulation.context.last_potential_energy)), "{:>12.4f}")
        self.add_field("<N>", lambda: float(simulation.batch.n_atoms.double().mean().item()), "{:>10.2f}")

    def remove_fields(self, pattern: str) -> None:
        for name in list(self.fields):
            names = name if isinstance(name, tuple) else (name,)
            if any(pattern in n for n in names):
                self.fields.pop(name)
