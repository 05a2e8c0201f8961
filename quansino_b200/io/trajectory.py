"""``TrajectoryObserver`` (``src/quansino/io/trajectory.py:15-57``): extended-XYZ frames.

The reference hands its ``Atoms`` to ASE's ``write_xyz``.  ASE is not a dependency of the GPU path, so the (small)
extended-XYZ writer lives here: per frame the atom count, a comment line ``Lattice="..." Properties=species:S:1:pos:R:3
[momenta:R:3] key=value ... pbc="T T T"`` and one ``symbol x y z`` line per atom with ASE's ``%16.8f`` columns.  A batched
simulation writes one frame per selected replica and call, tagged ``replica=<global id>`` (and ``step=`` when known)."""

from __future__ import annotations

from pathlib import Path
from typing import IO, Any

import numpy as np

from quansino_b200.io.core import TextObserver


def _fmt_value(v) -> str:
    if isinstance(v, (bool, np.bool_)):
        return "T" if v else "F"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if isinstance(v, (float, np.floating)):
        return repr(float(v))
    s = str(v)
    return f'"{s}"' if (" " in s or "=" in s) else s


def write_extxyz(file: IO, positions, cell, symbol: str, pbc=(True, True, True), info: dict | None = None, momenta=None) -> None:
    """One frame.  ``positions`` (n, 3), ``cell`` (3, 3) row-major lattice vectors."""
    positions = np.asarray(positions, dtype=float)
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    props = "species:S:1:pos:R:3" + (":momenta:R:3" if momenta is not None else "")
    lattice = " ".join(repr(float(x)) for x in cell.reshape(9))
    parts = [f'Lattice="{lattice}"', f"Properties={props}"]
    for k, v in (info or {}).items():
        parts.append(f"{k}={_fmt_value(v)}")
    parts.append('pbc="' + " ".join("T" if b else "F" for b in pbc) + '"')
    file.write(f"{len(positions)}\n")
    file.write(" ".join(parts) + "\n")
    cols = positions if momenta is None else np.hstack([positions, np.asarray(momenta, dtype=float)])
    for row in cols:
        file.write(f"{symbol:<2s} " + " ".join(f"{x:16.8f}" for x in row) + "\n")


def read_extxyz(file: IO | str | Path) -> list[dict[str, Any]]:
    """Frames written by :func:`write_extxyz` (or ASE's ``write_xyz`` for single-species orthorhombic systems)."""
    import re

    if isinstance(file, (str, Path)):
        with open(file, encoding="utf-8") as fh:
            return read_extxyz(fh)
    frames = []
    while True:
        line = file.readline()
        if not line.strip():
            break
        n = int(line)
        comment = file.readline().strip()
        info: dict[str, Any] = {}
        for m in re.finditer(r'(\w+)=("([^"]*)"|\S+)', comment):
            info[m.group(1)] = m.group(3) if m.group(3) is not None else m.group(2)
        rows = [file.readline().split() for _ in range(n)]
        data = np.array([[float(x) for x in r[1:]] for r in rows], dtype=float).reshape(n, -1)
        frame = {
            "symbols": [r[0] for r in rows],
            "positions": data[:, :3],
            "cell": np.array([float(x) for x in info.pop("Lattice").split()]).reshape(3, 3),
            "pbc": tuple(x == "T" for x in info.pop("pbc", "F F F").split()),
            "info": info,
        }
        if "momenta" in info.get("Properties", ""):
            frame["momenta"] = data[:, 3:6]
        frames.append(frame)
    return frames


class TrajectoryObserver(TextObserver):
    """``TrajectoryObserver(atoms, file, interval=1, mode="a", write_kwargs=None)``.

    ``atoms`` is the ``BatchedAtoms`` of the simulation (the driver refreshes it from the device before the call).
    ``write_kwargs``: ``replicas`` (iterable of local replica indices, default all), ``write_momenta`` (bool)."""

    def __init__(self, atoms, file: IO | Path | str, interval: int = 1, mode: str = "a", write_kwargs: dict[str, Any] | None = None,
                 **observer_kwargs: Any) -> None:  # fmt: skip
        super().__init__(file=file, interval=interval, mode=mode, **observer_kwargs)
        self.atoms = atoms
        self.write_kwargs = write_kwargs or {}
        self.simulation = None

    def attach_simulation(self, simulation=None, *_args: Any, **_kwargs: Any) -> None:
        self.simulation = simulation

    def __call__(self) -> None:
        sim = self.simulation
        if sim is not None:
            sim.download()
        atoms = self.atoms
        replicas = self.write_kwargs.get("replicas")
        replicas = range(atoms.n_replicas) if replicas is None else replicas
        energies = sim.context.last_potential_energy if sim is not None and hasattr(sim.context, "last_potential_energy") else None
        offset = sim.batch.replica_offset if sim is not None else 0
        for r in replicas:
            n = int(atoms.n_atoms[r])
            info: dict[str, Any] = {"replica": offset + int(r)}
            if sim is not None:
                info["step"] = int(sim.step_count)
            if energies is not None:
                info["energy"] = float(energies[r])
            mom = atoms.momenta[r, :n] if (self.write_kwargs.get("write_momenta") and atoms.momenta is not None) else None
            write_extxyz(self._file, atoms.positions[r, :n], atoms.cell[r], atoms.symbol, atoms.pbc, info, mom)
        self._file.flush()
