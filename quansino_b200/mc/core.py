"""The Monte Carlo driver (``src/quansino/mc/driver.py:27-176`` and ``src/quansino/mc/core.py:36-411``) over a batch
of replicas.

The host keeps quansino's loop -- ``run -> irun -> step``, move registration, observers, ``step_count`` -- while the
body of ``MonteCarlo.step`` (``mc/core.py:353-384``: select -> propose -> energy -> criteria -> save / revert, once per
cycle) is ONE kernel launch that performs ``max_cycles`` attempts for every replica.

Deviations from the reference, all explicit:

* ``step()`` still is a generator, but it yields the names of the moves available in this step once per launch, not
  one name per cycle (the per-cycle selection happens on the device, independently per replica);
* ``seed`` keys a counter-based Philox stream (DESIGN.md "Random stream") instead of PCG64;
* user-defined Python hooks and everything still listed as "next" in DESIGN.md section 0 raise
  ``NotImplementedError`` -- nothing falls back to the CPU.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Any, Callable, ClassVar
from warnings import warn

import numpy as np
import torch

from quansino_b200 import _lib
from quansino_b200.batch import BatchedAtoms, ReplicaBatch
from quansino_b200.mc.contexts import Context
from quansino_b200.mc.criteria import BaseCriteria
from quansino_b200.moves.core import BaseMove
from quansino_b200.utils.moves import MoveStorage


class MonteCarlo:
    """Batched counterpart of ``quansino.mc.core.MonteCarlo``.

    Parameters
    ----------
    atoms
        ``BatchedAtoms``, or one ASE-like ``Atoms`` together with ``replicas``, or a list of ``Atoms``.
    max_cycles
        Attempts per step and replica.
    seed
        Key of the random stream.
    replicas, calc, device, replica_offset
        Batch construction: number of copies of a single ``Atoms``, the potential description
        (``quansino_b200.calculators``), the CUDA device, and the global id of the first local replica (sharding).
    steps_per_launch
        ``run`` fuses up to this many steps into one kernel launch when no observer needs the state in between.
    resync_interval
        Every this many steps the caches are re-primed with a full energy evaluation (bounds the drift of the
        incrementally updated sigma_1 / energy); 0 disables.
    trace
        Keep the per-attempt trace of the last launch in ``last_trace`` (parity tests).
    """

    def __class_getitem__(cls, item):
        """``Isobaric[Move, Criteria](...)``, ``CellMove[OperationType, ContextType]``: the reference's classes are ``Generic``; the
        parameters carry no run-time meaning here either."""
        return cls


    default_criteria: ClassVar[dict[type, type]] = {BaseMove: BaseCriteria}
    default_context: ClassVar[type] = Context
    needs_momenta: ClassVar[bool] = False

    def __init__(
        self,
        atoms,
        max_cycles: int = 1,
        seed: int | None = None,
        replicas: int | None = None,
        calc=None,
        device: str | torch.device = "cuda",
        replica_offset: int = 0,
        steps_per_launch: int = 1,
        resync_interval: int = 100,
        trace: bool = False,
        logfile=None,
        trajectory=None,
        restart_file=None,
        logging_interval: int = 1,
        logging_mode: str = "a",
        neighbour_rows: bool | None = None,
    ) -> None:
        self.moves: dict[str, MoveStorage] = {}
        self.max_cycles = max_cycles
        self.acceptance_rate: np.ndarray | float = 0.0
        self.move_history: list[tuple[str, Any]] = []
        self._seed = int(seed) if seed is not None else int.from_bytes(os.urandom(8), "little")  # 0 is a valid seed (mc/driver.py:77-78)
        self.logging_interval = logging_interval
        self.logging_mode = logging_mode
        self.step_count = 0
        self.max_steps = 0
        self.attempt_counter = 0
        self.steps_per_launch = max(1, int(steps_per_launch))
        self.resync_interval = int(resync_interval)
        self.trace = bool(trace)
        self.last_trace: dict[str, np.ndarray] | None = None
        self.observers: dict[str, tuple[Callable[[], Any], int]] = {}

        if not isinstance(atoms, BatchedAtoms):
            atoms = BatchedAtoms.from_ase(atoms, replicas=replicas if not isinstance(atoms, (list, tuple)) else None, calc=calc)
        self.atoms = atoms
        self.batch = ReplicaBatch(
            atoms, calc=calc, device=device, replica_offset=replica_offset, with_momenta=self.needs_momenta, neighbour_rows=neighbour_rows
        )
        self.context = self.default_context(atoms, self.batch, self._seed)
        # default observers (mc/driver.py:42-63, 196-262): created from a path / stream, or attached as given
        self.default_logger = self.default_trajectory = self.default_restart = None
        if logfile is not None:
            from quansino_b200.io import Logger

            self.default_logger = logfile if isinstance(logfile, Logger) else Logger(logfile, logging_interval, logging_mode)
            self.default_logger.add_mc_fields(self)
            self.attach(self.default_logger, name="default_logger")
        if trajectory is not None:
            from quansino_b200.io import TrajectoryObserver

            self.default_trajectory = (
                trajectory if isinstance(trajectory, TrajectoryObserver) else TrajectoryObserver(atoms, trajectory, logging_interval, logging_mode)
            )
            self.attach(self.default_trajectory, name="default_trajectory")
        if restart_file is not None:
            from quansino_b200.io import RestartObserver

            self.default_restart = (
                restart_file if isinstance(restart_file, RestartObserver) else RestartObserver(self, restart_file, logging_interval, logging_mode)
            )
            self.attach(self.default_restart, name="default_restart")
        self._lowered: tuple | None = None
        self._slots: dict[str, int] = {}
        self._primed = False
        self._counters_before: torch.Tensor | None = None
        self._cycles_last_launch = 0

    # -- move registration (mc/core.py:133-184) -------------------------------------------------------
    def add_move(self, move, criteria=None, name: str = "default", interval: int = 1, probability: float = 1.0, minimum_count: int = 0) -> None:
        forced_moves_count = sum(self.moves[n].minimum_count for n in self.moves)
        if forced_moves_count + minimum_count > self.max_cycles:
            raise ValueError("The number of forced moves exceeds the number of cycles.")
        if criteria is None:
            for move_type in self.default_criteria:
                if isinstance(move, move_type):
                    criteria = self.default_criteria[move_type]()
                    break
        if criteria is None:
            raise ValueError(f"No criteria provided, and no default criteria found for move type {type(move)}.")
        # limits of the device path are reported where the move is registered, not at the first launch: the sweep kernels keep a
        # replica in shared memory (<= 1024 atoms); only Hamiltonian moves and ForceBias run on larger replicas
        if getattr(move, "move_type", None) in (_lib.MOVE_DISPLACEMENT, _lib.MOVE_CELL, _lib.MOVE_EXCHANGE) and self.batch.n_max > 1024:
            raise NotImplementedError(
                f"{type(move).__name__}: the sweep kernels hold a replica in shared memory, at most 1024 atoms per replica "
                f"(this batch has n_max = {self.batch.n_max}); Hamiltonian moves and ForceBias have no such limit"
            )
        self.moves[name] = MoveStorage(move=move, criteria=criteria, interval=interval, probability=probability, minimum_count=minimum_count)
        self._lowered = None

    # -- lowering -----------------------------------------------------------------------------------
    _criteria_for: ClassVar[dict[int, tuple[str, ...]]] = {
        _lib.MOVE_DISPLACEMENT: ("CanonicalCriteria",),
        _lib.MOVE_CELL: ("IsobaricCriteria", "IsotensionCriteria"),
        _lib.MOVE_EXCHANGE: ("GrandCanonicalCriteria",),
        _lib.MOVE_HMC: ("HamiltonianCanonicalCriteria",),
    }

    def _lower(self, names: list[str]):
        """POD move table for the given (available) move names; label tables are created once per move.  A composite
        move occupies a chain of consecutive entries; ``self._slots[name]`` is the entry of its head (counters, trace)."""
        # The table is lowered afresh at EVERY launch: the reference re-reads step sizes, operations, probabilities, Verlet
        # parameters and labels on every call, so changes made between runs must reach the kernel (a few small structs).
        # Only the device label tables persist (they are the labels' state once exchange moves re-label them).
        n_atoms = None  # fetched (one device sync) only when a label table has to be created
        any_exchange = any(self.moves[n].move.move_type == _lib.MOVE_EXCHANGE for n in names)
        # molecular exchange (several atoms per insertion, or TranslationRotation placement): every move of the table selects
        # label GROUPS, coded as runs of equal labels (include/quansino_b200.h: chain bit 3)
        molecular = any_exchange and (
            len(np.atleast_2d(self.batch.exchange_mol)) > 1
            or any(self.moves[n].move.move_type == _lib.MOVE_EXCHANGE and getattr(self.moves[n].move.operation, "op_code", -1) == _lib.OP_TRANSLATION_ROTATION for n in names)
            or any(self.moves[n].move.move_type == _lib.MOVE_EXCHANGE and len(getattr(self.moves[n].move, "moves", ())) > 1 for n in names)
            or any(self._has_label_groups(self.moves[n].move) for n in names)
        )  # fmt: skip  (a CompositeExchangeMove gives the atoms of one insertion ONE label: groups, too; and so are label groups
        # the caller defines himself -- `where(labels == label)` displaces / deletes all their atoms, moves/exchange.py:127)
        entries, keep, slots = [], [], {}
        cell_criteria = {type(self.moves[n].criteria).__name__ for n in names if self.moves[n].move.move_type == _lib.MOVE_CELL}
        if len(cell_criteria) > 1:
            raise NotImplementedError("cell moves judged by different criteria in one step are not available on the GPU path")
        # IsotensionCriteria lives in the general-cell kernels (csrc/sweep_tri.cuh), and so do the deformations that leave
        # general cells behind; a batch that was switched to them stays there (its cells may be triclinic by now)
        self.batch.isotension = cell_criteria == {"IsotensionCriteria"}
        if self.batch.isotension or any(
            getattr(getattr(self.moves[n].move, "operation", None), "op_code", -1) in (_lib.OP_ANISOTROPIC, _lib.OP_SHAPE) for n in names
        ):
            if not self.batch.triclinic:
                self.batch.triclinic = True
                self.batch.drop_rows()
        for n in names:
            st = self.moves[n]
            st.criteria._check_supported()
            chain = st.move.lower_chain() if hasattr(st.move, "lower_chain") else [st.move.lower()]
            m = chain[0]
            judged_as = _lib.MOVE_CELL if any(c.type == _lib.MOVE_CELL for c in chain) else m.type  # a hybrid changes the volume
            if type(st.criteria).__name__ not in self._criteria_for[judged_as]:
                raise NotImplementedError(
                    f"{type(st.criteria).__name__} with {type(st.move).__name__} is not a combination of the GPU path"
                )
            m.probability = float(st.probability)
            m.minimum_count = int(st.minimum_count)  # forced occurrences per step (mc/core.py:331-342), placed by the kernel
            if m.minimum_count and m.type == _lib.MOVE_HMC:
                raise NotImplementedError("forced Hamiltonian moves are not available on the GPU path")
            if m.type in (_lib.MOVE_DISPLACEMENT, _lib.MOVE_EXCHANGE):
                version = self._labels_version(st.move)
                if hasattr(st, "_label_tensor") and getattr(st, "_label_version", version) != version:
                    del st._label_tensor  # set_labels() was called since the table was uploaded
                st._label_version = version
                if n_atoms is None and (not hasattr(st, "_label_tensor") or (st._label_tensor is None and any_exchange)):
                    n_atoms = self.batch.n_atoms.cpu().numpy()
                make_table = st.move.run_labels if molecular else st.move.validated_labels
                if not hasattr(st, "_label_tensor"):
                    if any_exchange or not st.move.is_identity(n_atoms):
                        lab = make_table(self.batch.R, self.batch.n_max, n_atoms)
                        st._label_tensor = torch.from_numpy(lab).to(self.batch.device)
                    else:
                        st._label_tensor = None
                elif st._label_tensor is None and any_exchange:
                    lab = make_table(self.batch.R, self.batch.n_max, n_atoms)
                    st._label_tensor = torch.from_numpy(lab).to(self.batch.device)
                grouped = bool(getattr(st.move, "_grouped", False) or getattr(getattr(st.move, "moves", [None])[0], "_grouped", False))
                # rotations and relocations of displacement moves live in the label-group path
                grouped = grouped or any(c.type == _lib.MOVE_DISPLACEMENT and c.op in (_lib.OP_ROTATION, _lib.OP_TRANSLATION, _lib.OP_TRANSLATION_ROTATION) for c in chain)
                if molecular:
                    if len(chain) > 1 or (chain[0].repeat > 1 and m.type != _lib.MOVE_EXCHANGE) or m.n_extra_ops:
                        raise NotImplementedError("composite moves next to a molecular exchange move or label groups are not available on the GPU path")
                    m.chain = 8  # run-coded label groups
                    grouped = False
                if grouped and (any_exchange or len(chain) > 1 or chain[0].repeat > 1 or m.type != _lib.MOVE_DISPLACEMENT or m.n_extra_ops):
                    raise NotImplementedError(
                        "label groups (several atoms per label, or labels not increasing with the atom index) cannot be combined "
                        "with exchange moves or composite moves on the GPU path (SURVEY.md section 8f)"
                    )
                if grouped:
                    m.chain = 4  # labels are dense group ranks (include/quansino_b200.h)
                for sub in chain:  # the displacement sub-moves of a composite share one label table
                    if sub.type != _lib.MOVE_CELL:
                        sub.labels = st._label_tensor.data_ptr() if st._label_tensor is not None else None
                keep.append(st._label_tensor)
            slots[n] = len(entries)
            entries.extend(chain)
        if len(entries) > _lib.QMC_MAX_MOVES:
            raise NotImplementedError(f"at most {_lib.QMC_MAX_MOVES} move-table entries per step on the GPU path")
        arr = (_lib.Move * len(entries))()
        for i, m in enumerate(entries):
            arr[i] = m
        arr[0].max_cycles = int(self.max_cycles)
        self._slots = slots
        self._lowered = (tuple(names), arr, keep)  # keeps the ctypes array and the label tensors alive during the launch
        return arr, keep

    @staticmethod
    def _has_label_groups(move) -> bool:
        """Whether a displacement / exchange move selects GROUPS: a non-negative label carried by several atoms, or a rigid-body
        operation (Rotation / Translation / TranslationRotation of a displacement move)."""
        for m in getattr(move, "moves", [move]):
            if getattr(m, "move_type", None) not in (_lib.MOVE_DISPLACEMENT, _lib.MOVE_EXCHANGE) or not hasattr(m, "labels"):
                continue
            if m.move_type == _lib.MOVE_DISPLACEMENT and getattr(m.operation, "op_code", -1) in (_lib.OP_ROTATION, _lib.OP_TRANSLATION, _lib.OP_TRANSLATION_ROTATION):
                return True
            lab = np.atleast_2d(np.asarray(m.labels))
            for row in lab:
                v = row[row >= 0]
                if len(np.unique(v)) < len(v):
                    return True
        return False

    @staticmethod
    def _labels_version(move) -> tuple:
        subs = getattr(move, "moves", None) or [move]
        return tuple(getattr(m, "_labels_version", 0) for m in subs)

    def labels(self, name: str) -> np.ndarray | None:
        """Current label table [R][n_max] of a displacement / exchange move (None = arange)."""
        t = getattr(self.moves[name], "_label_tensor", None)
        return None if t is None else t.cpu().numpy()

    # -- driver loop (mc/driver.py:91-132) -------------------------------------------------------------
    def attach(self, function: Callable[[], Any], interval: int | None = None, name: str | None = None) -> None:
        """Attach an observer called every ``interval`` steps (``interval < 0``: once, at step ``abs(interval)``).
        ``function`` is a plain callable or a ``quansino_b200.io.Observer`` (its own ``interval`` is used)."""
        from quansino_b200.io.core import Observer

        if isinstance(function, Observer):
            if interval is not None:
                function.interval = int(interval)
            function.attach_simulation(self)
            entry = (function, function)
        else:
            entry = (function, 1 if interval is None else int(interval))
        self.observers[name or f"observer_{len(self.observers)}"] = entry

    @staticmethod
    def _interval_of(entry) -> int:
        return int(entry[1].interval) if hasattr(entry[1], "interval") else int(entry[1])

    def close(self) -> None:
        """Close the files of the attached observers (``mc/driver.py:264-266``)."""
        for function, _ in self.observers.values():
            if hasattr(function, "close"):
                function.close()

    def call_observers(self) -> None:
        for entry in self.observers.values():
            function, interval = entry[0], self._interval_of(entry)
            if (interval > 0 and self.step_count % interval == 0) or (interval < 0 and self.step_count == abs(interval)):
                function()

    def validate_simulation(self) -> None:
        if len(self.moves) == 0:
            warn("No moves have been added to the Monte Carlo simulation. Please add moves using the `add_move` method.", UserWarning, 2)
        if not self._primed:
            self.batch.energy_full()
            self._primed = True

    def converged(self) -> bool:
        return self.step_count >= self.max_steps

    def _steps_this_launch(self) -> int:
        k = min(self.steps_per_launch, self.max_steps - self.step_count)
        if any(st.interval != 1 for st in self.moves.values()):
            return 1
        for entry in self.observers.values():
            interval = self._interval_of(entry)
            if interval > 0:
                k = min(k, interval - self.step_count % interval)
            elif interval < 0 and abs(interval) > self.step_count:
                k = min(k, abs(interval) - self.step_count)
        if self.resync_interval > 0:
            k = min(k, self.resync_interval - self.step_count % self.resync_interval)
        return max(1, k)

    def irun(self, steps: int = 100_000_000):
        self.validate_simulation()
        self.max_steps = self.step_count + steps
        if self.step_count == 0:
            self.call_observers()
        while not self.converged():
            k = self._steps_this_launch()
            yield self.step(k)
            self.step_count += k
            if self.resync_interval > 0 and self.step_count % self.resync_interval == 0:
                self.batch.energy_full()
            self.call_observers()

    def srun(self, steps: int = 100_000_000):
        for step in self.irun(steps):
            for _ in step:
                pass
            yield step

    def run(self, steps: int = 100_000_000) -> None:
        for step in self.irun(steps):
            for _ in step:
                pass
        self.batch.raise_on_status()

    def yield_moves(self):
        """Names of the moves available at this step (``mc/core.py:326-333``); the per-cycle choice is made on the
        device from ``u_select`` with the normalised probabilities."""
        return [name for name in self.moves if self.step_count % self.moves[name].interval == 0]

    def step(self, n_steps: int = 1):
        """One (or ``n_steps`` fused) Monte Carlo step(s): ``n_steps * max_cycles`` attempts per replica in one launch."""
        self.move_history = []
        names = self.yield_moves()
        if not names:
            return
        if not self._primed:
            self.validate_simulation()
        for name in names:
            yield name
        n_attempts = int(n_steps) * int(self.max_cycles)
        self._launch(names, n_attempts)
        self.attempt_counter += n_attempts
        self._cycles_last_launch = n_attempts
        self.move_history = [(name, None) for name in names]

    def _launch(self, names: list[str], n_attempts: int) -> None:
        moves, _keep = self._lower(names)
        b = self.batch
        self._counters_before = b.counters.clone()
        tr = None
        bufs = None
        if self.trace:
            R = b.R
            bufs = {
                "move": torch.zeros((R, n_attempts), dtype=torch.int8, device=b.device),
                "accepted": torch.zeros((R, n_attempts), dtype=torch.int8, device=b.device),
                "energy": torch.zeros((R, n_attempts), dtype=torch.float64, device=b.device),
                "delta": torch.zeros((R, n_attempts), dtype=torch.float64, device=b.device),
                "moved": torch.zeros((R, n_attempts), dtype=torch.int32, device=b.device),
            }
            tr = _lib.Trace(*[bufs[k].data_ptr() for k in ("move", "accepted", "energy", "delta", "moved")])
        use_rows = self._use_rows(moves, n_attempts)
        if use_rows:
            b.ensure_rows()
        else:
            b.rows_valid = False  # whatever runs now moves atoms without looking at the rows
        with torch.cuda.device(b.device):
            bs = b.struct(with_rows=use_rows)
            if moves[0].type == _lib.MOVE_HMC:
                if len(names) != 1:
                    raise NotImplementedError("Hamiltonian moves cannot be mixed with other moves in one step on the GPU path")
                ws = b.hmc_workspace(bs)
                rc = b.lib.qmc_hmc(
                    C.byref(b.potential), C.byref(bs), C.byref(moves[0]), self._seed, self.attempt_counter, n_attempts,
                    C.byref(tr) if tr is not None else None, ws.data_ptr(), ws.numel(), b.stream(),
                )  # fmt: skip
            else:
                rc = b.lib.qmc_sweep(
                    C.byref(b.potential), C.byref(bs), moves, len(moves), self._seed, self.attempt_counter, n_attempts,
                    C.byref(tr) if tr is not None else None, b.stream(),
                )  # fmt: skip
        _lib.check(rc)
        if any(m.type == _lib.MOVE_CELL for m in moves):
            b.cells_uniform = False  # accepted cell moves make the cells replica-specific
        if bufs is not None:
            self.last_trace = {k: v.cpu().numpy() for k, v in bufs.items()}

    rows_min_atoms: ClassVar[int] = 256  # auto mode: neighbour rows from this capacity on (profiles/r2_a_rows.md)

    def _use_rows(self, moves, n_attempts: int) -> bool:
        """Whether this launch runs the row-based sweep kernel (``csrc/sweep_rows.cuh``): tables of single-atom displacement
        moves (Ball / Box / Sphere, steps within half the skin) and cell moves, on batches whose cells leave room for cutoff +
        skin.  ``neighbour_rows=True`` / ``False`` (or ``QMC_FORCE_KERNEL``) decide; by default the rows are used when they are
        already built or when the launch is long enough to pay for building them."""
        b = self.batch
        force = os.environ.get("QMC_FORCE_KERNEL")
        want = b.want_rows if force is None else (True if force == "rows" else False)
        if want is False or not b.rows_possible():
            return False
        skin = b.nbr_skin or _lib.NBR_SKIN_DEFAULT
        for m in moves:
            if m.type == _lib.MOVE_CELL:
                continue
            if m.type != _lib.MOVE_DISPLACEMENT or m.op not in (_lib.OP_BALL, _lib.OP_BOX, _lib.OP_SPHERE):
                return False
            if m.chain or m.repeat > 1 or m.n_extra_ops or m.minimum_count:
                return False
            if abs(m.step) * (3.0**0.5 if m.op == _lib.OP_BOX else 1.0) * (1.0 + 1e-9) > 0.5 * skin:
                return False
        if want:
            return True
        return b.n_max >= self.rows_min_atoms

    # -- results ------------------------------------------------------------------------------------
    def counters(self) -> np.ndarray:
        """[R][move slot][attempted, accepted, failed], accumulated since construction."""
        return self.batch.counters.cpu().numpy()

    def update_acceptance_rate(self) -> np.ndarray:
        """``acceptance_rate`` of the last launch per replica: accepted / cycles (failed moves count as rejected,
        ``mc/core.py:382-384``)."""
        if self._counters_before is None or self._cycles_last_launch == 0:
            self.acceptance_rate = np.zeros(self.batch.R)
        else:
            d = (self.batch.counters - self._counters_before)[:, :, 1].sum(dim=1).cpu().numpy()
            self.acceptance_rate = d / float(self._cycles_last_launch)
        return self.acceptance_rate

    def download(self) -> BatchedAtoms:
        """Copy the device state back into ``self.atoms`` and return it."""
        self.batch.download(self.atoms)
        return self.atoms

    def to_dict(self) -> dict[str, Any]:
        """Everything a restart needs (``mc/driver.py:177-194``, ``mc/core.py:244-309``).  The reference stores
        ``rng.bit_generator.state``; here ``seed`` and ``attempt_counter`` ARE the state of the counter-based stream.  The
        incrementally tracked energies and EMT caches are stored too, so that a restart continues bit-identically."""
        b = self.batch
        self.download()
        state = {
            "positions": self.atoms.positions, "cell": self.atoms.cell, "n_atoms": self.atoms.n_atoms, "symbol": self.atoms.symbol,
            "pbc": list(self.atoms.pbc), "energy": b.energy.cpu().numpy(), "sigma1": b.sigma1.cpu().numpy(), "eatom": b.eatom.cpu().numpy(),
            "temperature": b.temperature.cpu().numpy(), "counters": b.counters.cpu().numpy(), "n_exchange": b.n_exchange.cpu().numpy(),
            "replica_offset": b.replica_offset, "cells_uniform": bool(b.cells_uniform),
        }  # fmt: skip
        if b.momenta is not None:
            state["momenta"] = self.atoms.momenta
            state["kinetic"] = b.kinetic.cpu().numpy()
        labels = {name: self.labels(name) for name in self.moves if getattr(self.moves[name], "_label_tensor", None) is not None}
        return {
            "name": self.__class__.__name__,
            "kwargs": {"seed": self._seed, "max_cycles": self.max_cycles, "logging_interval": self.logging_interval,
                       "logging_mode": self.logging_mode},
            "attributes": {"step_count": self.step_count, "attempt_counter": self.attempt_counter, "primed": self._primed},
            "context": self.context.to_dict(),
            "moves": {name: st.to_dict() for name, st in self.moves.items()},
            "move_order": list(self.moves),  # the files are written with sorted keys; the ORDER of the moves is part of the state
            "state": state,
            "labels": labels,
        }  # fmt: skip

    def load_state(self, data: dict[str, Any]) -> None:
        """Continue a run from ``to_dict()`` / ``quansino_b200.io.read_restart``.  The simulation must have been constructed
        like the one that wrote the restart (same class, moves in the same order, same batch shape)."""
        if data.get("name") != self.__class__.__name__:
            raise ValueError(f"restart data of a {data.get('name')} cannot be loaded into a {self.__class__.__name__}")
        if list(data.get("move_order", data.get("moves", {}))) != list(self.moves):
            raise ValueError("the restart data was written with a different set of moves")
        st, b = data["state"], self.batch
        pos = np.asarray(st["positions"], dtype=np.float64)
        if pos.shape != self.atoms.positions.shape:
            raise ValueError(f"restart batch shape {pos.shape} differs from this simulation's {self.atoms.positions.shape}")
        if int(data["kwargs"]["seed"]) != self._seed:
            self._seed = int(data["kwargs"]["seed"])
            self.context.seed = self._seed
        self.atoms.positions[...] = pos
        self.atoms.cell[...] = np.asarray(st["cell"], dtype=np.float64)
        self.atoms.n_atoms[...] = np.asarray(st["n_atoms"], dtype=np.int32)
        if "momenta" in st and b.momenta is not None:
            self.atoms.momenta = np.array(st["momenta"], dtype=np.float64)
        b.upload(self.atoms)
        dev = b.device
        for name, dtype in (("energy", np.float64), ("sigma1", np.float64), ("eatom", np.float64), ("temperature", np.float64),
                            ("counters", np.int64), ("n_exchange", np.int32)):  # fmt: skip
            getattr(b, name).copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(st[name], dtype=dtype))).to(dev))
        if "kinetic" in st and b.kinetic is not None:
            b.kinetic.copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(st["kinetic"], dtype=np.float64))).to(dev))
        b.cells_uniform = bool(st.get("cells_uniform", b.cells_uniform)) and b.cells_uniform
        self.step_count = int(data["attributes"]["step_count"])
        self.attempt_counter = int(data["attributes"]["attempt_counter"])
        self._primed = bool(data["attributes"].get("primed", True))
        self._lowered = None
        for name, table in data.get("labels", {}).items():
            self.moves[name]._label_tensor = torch.from_numpy(np.ascontiguousarray(np.asarray(table, dtype=np.int32))).to(dev)

    todict = to_dict
