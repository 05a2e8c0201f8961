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
