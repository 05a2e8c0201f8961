"""``RestartObserver`` (``src/quansino/io/restart.py:16-71``): the simulation's ``to_dict()`` as JSON, rewritten in place.

The reference serialises through ``ase.io.jsonio.write_json``; ASE is not a dependency here, so numpy arrays are encoded
the way ASE does it -- ``{"__ndarray__": [shape, dtype, flat list]}`` -- by a small encoder, and ``read_restart`` decodes
them.  The random stream of the GPU path is counter-based: ``seed`` + ``attempt_counter`` in the dictionary ARE the RNG
state (the reference stores ``rng.bit_generator.state``), so a restarted run continues bit-identically
(``tests/test_io.py``, the counterpart of ``tests/mc/test_canonical.py:268-306``)."""

from __future__ import annotations

import json
from pathlib import Path
from typing import IO, Any

import numpy as np

from quansino_b200.io.core import TextObserver


class _Encoder(json.JSONEncoder):
    def default(self, o):
        if isinstance(o, np.ndarray):
            flat = o.ravel()
            data = flat.tolist() if o.dtype != np.complex128 else [[z.real, z.imag] for z in flat]
            return {"__ndarray__": [list(o.shape), str(o.dtype), data]}
        if isinstance(o, np.integer):
            return int(o)
        if isinstance(o, np.floating):
            return float(o)
        if isinstance(o, np.bool_):
            return bool(o)
        if hasattr(o, "to_dict"):
            return o.to_dict()
        if hasattr(o, "todict"):
            return o.todict()
        return repr(o)


def _decode(o):
    if isinstance(o, dict):
        if "__ndarray__" in o and len(o) == 1:
            shape, dtype, data = o["__ndarray__"]
            return np.array(data, dtype=dtype).reshape(shape)
        return {k: _decode(v) for k, v in o.items()}
    if isinstance(o, list):
        return [_decode(v) for v in o]
    return o


def write_json(file: IO, obj: Any, **kwargs: Any) -> None:
    data = obj.to_dict() if hasattr(obj, "to_dict") else obj
    json.dump(data, file, cls=_Encoder, **{"sort_keys": True, **kwargs})


def read_restart(file: IO | str | Path) -> dict[str, Any]:
    """The dictionary a ``RestartObserver`` wrote; feed it to ``MonteCarlo.load_state``."""
    if isinstance(file, (str, Path)):
        with open(file, encoding="utf-8") as fh:
            return _decode(json.load(fh))
    return _decode(json.load(file))


class RestartObserver(TextObserver):
    accept_stream: bool = False

    def __init__(self, simulation, file: IO | Path | str, interval: int = 1, mode: str = "a", write_kwargs: dict[str, Any] | None = None,
                 **observer_kwargs: Any) -> None:  # fmt: skip
        super().__init__(file=file, interval=interval, mode=mode, **observer_kwargs)
        self.simulation = simulation
        self.write_kwargs = write_kwargs or {}

    def __call__(self) -> None:
        self._file.seek(0)
        self._file.truncate()
        write_json(self._file, obj=self.simulation, **self.write_kwargs)
        self._file.flush()


def restore_simulation(data: dict[str, Any] | IO | str | Path, calc, device="cuda", **overrides: Any):
    """Rebuild a simulation from a restart dictionary / file: the driver class, its moves and criteria come back through the
    class registry (``quansino_b200.registry``), the batch from the stored state, and the run continues bit-identically.
    ``calc`` is the potential description (``quansino_b200.calculators``); it is not stored in the file, like the ASE
    calculator of the reference."""
    import numpy as np

    from quansino_b200.batch import BatchedAtoms
    from quansino_b200.registry import get_class
    from quansino_b200.utils.moves import MoveStorage

    if not isinstance(data, dict):
        data = read_restart(data)
    cls = get_class(data["name"])
    st, ctx, kw0 = data["state"], data.get("context", {}), data.get("kwargs", {})
    atoms = BatchedAtoms(np.array(st["positions"]), np.array(st["cell"]), np.array(st["n_atoms"], dtype=np.int32), st["symbol"],
                         tuple(bool(b) for b in st["pbc"]), calc, momenta=np.array(st["momenta"]) if "momenta" in st else None)  # fmt: skip
    kw: dict[str, Any] = {"seed": int(kw0["seed"]), "device": device, "replica_offset": int(st.get("replica_offset", 0))}
    names = {c.__name__ for c in cls.__mro__}
    if "ForceBias" in names:
        kw["delta"] = float(kw0["delta"])
    else:
        kw["max_cycles"] = int(kw0["max_cycles"])
    if "Canonical" in names or "ForceBias" in names:
        kw["temperature"] = np.array(st["temperature"])
    if "Isobaric" in names:
        kw["pressure"] = float(ctx["pressure"])
    if "Isotension" in names and "external_stress" in ctx:
        kw["external_stress"] = np.array(ctx["external_stress"], dtype=float).reshape(3, 3)
    if "GrandCanonical" in names:
        kw["chemical_potential"] = float(ctx["chemical_potential"])
        kw["number_of_exchange_particles"] = np.array(st["n_exchange"])
        ex = ctx.get("exchange_atoms")
        if ex is not None:
            kw["exchange_atoms"] = (ex[0], tuple(ex[1])) if isinstance(ex, (list, tuple)) else ex
    kw.update(overrides)
    sim = cls(atoms, **kw)
    if "GrandCanonical" in names and "accessible_volume" in ctx:
        sim.accessible_volume = float(ctx["accessible_volume"])
    stored_moves = data.get("moves", {})
    for name in data.get("move_order", list(stored_moves)):  # (JSON objects come back key-sorted: the list keeps the table order)
        stored = stored_moves[name]
        ms = MoveStorage.from_dict(stored)
        sim.add_move(ms.move, criteria=ms.criteria, name=name, interval=ms.interval, probability=ms.probability, minimum_count=ms.minimum_count)
    sim.load_state(data)
    return sim
