"""Replica batches: the host-side container (``BatchedAtoms``) and the device-resident state (``ReplicaBatch``).

The reference drives ONE ``ase.Atoms`` per simulation and reserves the names ``MultiDriver`` / ``MultiContexts``
for a multi-system API it never filled in (``src/quansino/mc/driver.py:349-371``, ``src/quansino/mc/contexts.py:393-436``).
``BatchedAtoms`` is the many-replica ``Atoms``; ``ReplicaBatch`` is what ``Context`` (``mc/contexts.py:20-96``)
owns per replica -- positions, cell, ``last_potential_energy``, calculator caches -- laid out for the kernels:

    pos     [R][3][n_max] f64    x-row, y-row, z-row per replica (coalesced per-replica loads)
    cell    [R][9]        f64    row-major, diagonal
    n_atoms [R]           i32
    energy  [R]           f64    last_potential_energy
    sigma1 / eatom [R][n_max] f64   EMT caches (``calc.results`` in the reference)
    counters [R][4][3]    i64    attempted / accepted / failed per move slot

torch is used for device memory and streams only.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from quansino_b200 import _lib
from quansino_b200.calculators import ATOMIC_MASSES, Calculator


@dataclass
class BatchedAtoms:
    """``R`` replicas of a single-species system, padded to ``n_max`` atoms."""

    positions: np.ndarray  # (R, n_max, 3)
    cell: np.ndarray  # (R, 3, 3)
    n_atoms: np.ndarray  # (R,) int32
    symbol: str = "Cu"
    pbc: tuple = (True, True, True)
    calc: Calculator | None = None
    momenta: np.ndarray | None = None  # (R, n_max, 3)
    info: dict = field(default_factory=dict)

    def __post_init__(self):
        self.positions = np.ascontiguousarray(self.positions, dtype=np.float64)
        R, n_max, _ = self.positions.shape
        cell = np.asarray(self.cell, dtype=np.float64)
        if cell.shape == (3,):
            cell = np.diag(cell)
        if cell.shape == (3, 3):
            cell = np.broadcast_to(cell, (R, 3, 3))
        self.cell = np.array(cell, dtype=np.float64, order="C", copy=True)
        self.n_atoms = np.array(np.broadcast_to(np.asarray(self.n_atoms, dtype=np.int32), (R,)), order="C", copy=True)
        self.pbc = tuple(bool(b) for b in np.broadcast_to(np.asarray(self.pbc), (3,)))
        if self.momenta is not None:
            self.momenta = np.ascontiguousarray(self.momenta, dtype=np.float64)

    @classmethod
    def replicate(cls, positions, cell, replicas: int, symbol="Cu", pbc=(True, True, True), calc=None, n_max=None):
        """``replicas`` copies of one configuration."""
        positions = np.asarray(positions, dtype=np.float64)
        n = len(positions)
        n_max = n if n_max is None else int(n_max)
        pos = np.zeros((replicas, n_max, 3))
        pos[:, :n] = positions
        return cls(pos, np.asarray(cell, dtype=np.float64), np.full(replicas, n, dtype=np.int32), symbol, pbc, calc)

    @classmethod
    def from_ase(cls, atoms, replicas: int | None = None, calc=None, n_max=None):
        """From one ASE-like ``Atoms`` (replicated) or a list of them (duck-typed: ``positions``, ``cell``,
        ``pbc``, ``get_chemical_symbols()``)."""
        lst = [atoms] * int(replicas) if replicas is not None else list(atoms)
        symbols = {s for a in lst for s in a.get_chemical_symbols()}
        if len(symbols) > 1:
            raise NotImplementedError(f"single-species systems only on the GPU path, got {sorted(symbols)}")
        n_max = max(len(a.positions) for a in lst) if n_max is None else int(n_max)
        pos = np.zeros((len(lst), n_max, 3))
        cells = np.zeros((len(lst), 3, 3))
        n = np.zeros(len(lst), dtype=np.int32)
        for r, a in enumerate(lst):
            n[r] = len(a.positions)
            pos[r, : n[r]] = a.positions
            cells[r] = np.asarray(a.cell)
        symbol = symbols.pop() if symbols else "X"
        return cls(pos, cells, n, symbol, tuple(lst[0].pbc), calc if calc is not None else getattr(lst[0], "calc", None))

    def __len__(self):
        return len(self.positions)

    @property
    def n_replicas(self) -> int:
        return self.positions.shape[0]

    @property
    def n_max(self) -> int:
        return self.positions.shape[1]

    def get_volume(self) -> np.ndarray:
        return np.abs(np.linalg.det(self.cell))

    def replica(self, r: int) -> np.ndarray:
        """Positions of replica ``r`` (n_atoms[r], 3)."""
        return self.positions[r, : self.n_atoms[r]]


def cell_heights(cell: np.ndarray) -> np.ndarray:
    """Perpendicular widths of (R, 3, 3) cells: w_k = |det| / |a_(k+1) x a_(k+2)| (what bounds the image search,
    ``ase/neighborlist.py``)."""
    cell = np.asarray(cell, dtype=np.float64).reshape(-1, 3, 3)
    vol = np.abs(np.linalg.det(cell))
    out = np.empty((len(cell), 3))
    for k in range(3):
        out[:, k] = vol / np.linalg.norm(np.cross(cell[:, (k + 1) % 3], cell[:, (k + 2) % 3]), axis=1)
    return out


def _is_diagonal(cell: np.ndarray) -> bool:
    off = np.array(cell, dtype=np.float64, copy=True).reshape(-1, 3, 3)
    off[:, [0, 1, 2], [0, 1, 2]] = 0.0
    return not np.any(off != 0.0)


def _check_cell(cell: np.ndarray, pbc, cutoff: float) -> bool:
    """Checks the cells against the image search of the kernels (two images per axis: every periodic height >= the cutoff);
    returns True if some cell is not diagonal (-> the general-cell kernels, ``csrc/sweep_tri.cuh``)."""
    if len(cell) == 0:
        return False
    if np.any(np.abs(np.linalg.det(cell)) <= 0.0):
        raise ValueError("singular cell")
    h = cell_heights(cell)
    for k in range(3):
        if pbc[k] and np.any(h[:, k] < cutoff * (1.0 + 1e-9)):
            raise NotImplementedError(
                f"periodic cell height {h[:, k].min():.3f} A along axis {k} is shorter than the neighbour cutoff "
                f"{cutoff:.3f} A: more than two images per axis are not available on the GPU path"
            )
    return not _is_diagonal(cell)


class ReplicaBatch:
    """Device-resident state of the replicas held by ONE GPU (``qmc_batch_t`` of ``include/quansino_b200.h``)."""

    def __init__(
        self,
        atoms: BatchedAtoms,
        calc: Calculator | None = None,
        temperature=298.15,
        device: str | torch.device = "cuda",
        replica_offset: int = 0,
        with_momenta: bool = False,
        neighbour_rows: bool | None = None,
    ):
        self.lib = _lib.load()
        self.calc = calc if calc is not None else atoms.calc
        if not isinstance(self.calc, Calculator):
            raise NotImplementedError(
                "the GPU path evaluates quansino_b200.calculators.EMT / LennardJones only; "
                f"got calculator {type(self.calc).__name__}"
            )
        if not torch.cuda.is_available():
            raise _lib.NoDeviceError("no CUDA device: quansino_b200 has no CPU fallback")
        self.device = torch.device(device)
        self.symbol = atoms.symbol
        self.pbc = atoms.pbc
        self.potential = self.calc.lower(atoms.symbol)
        self.R, self.n_max = atoms.n_replicas, atoms.n_max
        self.replica_offset = int(replica_offset)
        # general (triclinic) cells run through their own kernels; drivers also switch this on for moves that leave general
        # cells behind (AnisotropicDeformation, ShapeDeformation) or need the IsotensionCriteria
        self.triclinic = _check_cell(atoms.cell, self.pbc, self.calc.cutoff)
        self.isotension = False
        self.external_stress = np.zeros((3, 3))
        dev, f64 = self.device, torch.float64
        R, N = self.R, self.n_max
        self.pos = torch.zeros((R, 3, N), dtype=f64, device=dev)
        self.cell = torch.zeros((R, 9), dtype=f64, device=dev)
        self.n_atoms = torch.zeros(R, dtype=torch.int32, device=dev)
        self.energy = torch.full((R,), float("nan"), dtype=f64, device=dev)
        self.temperature = torch.zeros(R, dtype=f64, device=dev)
        self.sigma1 = torch.zeros((R, N), dtype=f64, device=dev)
        self.eatom = torch.zeros((R, N), dtype=f64, device=dev)
        self.momenta = torch.zeros((R, 3, N), dtype=f64, device=dev) if with_momenta else None
        self.kinetic = torch.full((R,), float("nan"), dtype=f64, device=dev) if with_momenta else None
        self.n_exchange = torch.zeros(R, dtype=torch.int32, device=dev)
        self.counters = torch.zeros((R, _lib.QMC_MAX_MOVES, 3), dtype=torch.int64, device=dev)
        self.status = torch.zeros(R, dtype=torch.int32, device=dev)
        self.pair_evals = torch.zeros(R, dtype=torch.int64, device=dev)
        self.mass = ATOMIC_MASSES.get(atoms.symbol, 1.0)
        self.pressure = 0.0
        self.chemical_potential = float("nan")
        self.accessible_volume = float(abs(np.linalg.det(atoms.cell[0]))) if R else 0.0
        self.exchange_mass = 0.0
        self.exchange_pos = (0.0, 0.0, 0.0)
        self.exchange_mol = np.zeros((1, 3))  # positions of the exchange molecule (one row: the single exchange atom)
        self.cells_uniform = False
        self.cell_diag = (0.0, 0.0, 0.0)
        # neighbour rows of the row-based sweep kernel (csrc/sweep_rows.cuh), allocated on first use (`ensure_rows`)
        self.want_rows = neighbour_rows
        self.nbr_rows = self.nbr_ref = self.nbr_state = None
        self.nbr_row_cap = 0
        self.nbr_skin = 0.0
        self.rows_valid = False  # host-side: False = the device flags nbr_state must be zeroed before the next row-based launch
        self.set_temperature(temperature)
        self.upload(atoms)

    # -- transfers ---------------------------------------------------------------------------------
    def upload(self, atoms: BatchedAtoms, non_blocking: bool = False) -> int:
        """Host -> device copy of positions, cells and atom counts; returns the bytes moved."""
        pos = torch.from_numpy(np.ascontiguousarray(atoms.positions.transpose(0, 2, 1)))
        cell = torch.from_numpy(atoms.cell.reshape(self.R, 9))
        n = torch.from_numpy(atoms.n_atoms)
        self.pos.copy_(pos, non_blocking=non_blocking)
        self.cell.copy_(cell, non_blocking=non_blocking)
        self.n_atoms.copy_(n, non_blocking=non_blocking)
        nbytes = pos.numel() * 8 + cell.numel() * 8 + n.numel() * 4
        self.rows_valid = False
        if self.R and not _is_diagonal(atoms.cell):
            self.triclinic = True
        diag = cell_heights(atoms.cell) if self.R else np.zeros((1, 3))  # (= the edges of diagonal cells)
        self._min_edge = tuple(float(diag[:, k].min()) for k in range(3))
        # one fixed cell shared by every replica: the kernels read it from their arguments (fast path)
        self.cells_uniform = bool(self.R > 0 and np.all(atoms.cell == atoms.cell[0]))
        self.cell_diag = tuple(float(atoms.cell[0, k, k]) for k in range(3)) if self.R else (0.0, 0.0, 0.0)
        if self.momenta is not None and atoms.momenta is not None:
            mom = torch.from_numpy(np.ascontiguousarray(atoms.momenta.transpose(0, 2, 1)))
            self.momenta.copy_(mom, non_blocking=non_blocking)
            nbytes += mom.numel() * 8
        return nbytes

    def download(self, atoms: BatchedAtoms) -> int:
        """Device -> host copy into ``atoms`` (positions, cell, n_atoms, momenta); returns the bytes moved."""
        pos = self.pos.cpu().numpy()
        atoms.positions[...] = pos.transpose(0, 2, 1)
        atoms.cell[...] = self.cell.cpu().numpy().reshape(self.R, 3, 3)
        atoms.n_atoms[...] = self.n_atoms.cpu().numpy()
        nbytes = pos.size * 8 + self.R * 72 + self.R * 4
        if self.momenta is not None:
            if atoms.momenta is None:
                atoms.momenta = np.zeros_like(atoms.positions)
            atoms.momenta[...] = self.momenta.cpu().numpy().transpose(0, 2, 1)
            nbytes += pos.size * 8
        return nbytes

    def positions(self) -> np.ndarray:
        """(R, n_max, 3) copy on the host."""
        return np.ascontiguousarray(self.pos.cpu().numpy().transpose(0, 2, 1))

    def set_temperature(self, temperature) -> None:
        t = np.broadcast_to(np.asarray(temperature, dtype=np.float64), (self.R,))
        if np.any(t == 0.0):
            raise ZeroDivisionError("float division by zero")  # what mc/criteria.py:108-110 raises at T = 0
        self.temperature.copy_(torch.from_numpy(np.array(t, order="C", copy=True)))

    # -- C struct ----------------------------------------------------------------------------------
    def struct(self, with_rows: bool = False) -> _lib.Batch:
        """``qmc_batch_t`` of the current state; the neighbour rows ride along only for launches that may use them."""
        b = _lib.Batch()
        b.n_replicas, b.n_max, b.replica_offset = self.R, self.n_max, self.replica_offset
        b.pbc = (C.c_int32 * 3)(*[int(x) for x in self.pbc])
        b.cells_uniform = int(self.cells_uniform)
        rows = with_rows and self.nbr_rows is not None
        b.nbr_row_cap = int(self.nbr_row_cap) if rows else 0
        b.nbr_rows = self.nbr_rows.data_ptr() if rows else None
        b.nbr_ref = self.nbr_ref.data_ptr() if rows else None
        b.nbr_state = self.nbr_state.data_ptr() if rows else None
        b.nbr_skin = float(self.nbr_skin)
        b.cell_diag = (C.c_double * 3)(*self.cell_diag)
        b.pos, b.cell, b.n_atoms = self.pos.data_ptr(), self.cell.data_ptr(), self.n_atoms.data_ptr()
        b.energy, b.temperature = self.energy.data_ptr(), self.temperature.data_ptr()
        b.sigma1, b.eatom = self.sigma1.data_ptr(), self.eatom.data_ptr()
        b.momenta = self.momenta.data_ptr() if self.momenta is not None else None
        b.kinetic = self.kinetic.data_ptr() if self.kinetic is not None else None
        b.n_exchange = self.n_exchange.data_ptr()
        b.counters, b.status, b.pair_evals = self.counters.data_ptr(), self.status.data_ptr(), self.pair_evals.data_ptr()
        b.mass, b.pressure, b.chemical_potential = self.mass, self.pressure, self.chemical_potential
        b.accessible_volume, b.exchange_mass = self.accessible_volume, self.exchange_mass
        b.exchange_pos = (C.c_double * 3)(*self.exchange_pos)
        mol = np.asarray(self.exchange_mol, dtype=np.float64).reshape(-1, 3)
        b.n_exchange_atoms = len(mol)
        b.exchange_mol = (C.c_double * 24)(*mol.reshape(-1), *([0.0] * (24 - mol.size)))
        b.triclinic, b.isotension = int(self.triclinic), int(self.isotension)
        b.external_stress = (C.c_double * 9)(*np.asarray(self.external_stress, dtype=np.float64).reshape(9))
        return b

    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # -- neighbour rows of the row-based sweep kernel ------------------------------------------------
    def rows_possible(self, skin: float | None = None) -> bool:
        """Whether this batch can carry neighbour rows: at most 1024 atoms per replica and every periodic edge at least
        cutoff + skin (checked against the cells known to the host; the kernel re-checks per replica)."""
        if self.n_max > 1024 or self.R == 0 or self.triclinic:
            return False
        skin = float(skin if skin is not None else (self.nbr_skin or _lib.NBR_SKIN_DEFAULT))
        need = self.calc.cutoff * (1.0 + 1e-9) + skin
        return all((not self.pbc[k]) or self._min_edge[k] >= need for k in range(3))

    def ensure_rows(self) -> None:
        """Allocate the neighbour rows (``qmc_batch_t.nbr_*``) on first use; the kernel builds them when ``nbr_state`` is 0."""
        if self.nbr_rows is None:
            words = int(self.lib.qmc_nbr_row_words(C.byref(self.potential), self.n_max))
            if words < 0:
                _lib.check(words)
            self.nbr_row_cap = words
            self.nbr_rows = torch.zeros((self.R, self.n_max, words), dtype=torch.int32, device=self.device)
            self.nbr_ref = torch.zeros((self.R, 8), dtype=torch.float64, device=self.device)
            self.nbr_state = torch.zeros(self.R, dtype=torch.int32, device=self.device)
            self.rows_valid = True  # (the zeroed state says "build")
        elif not self.rows_valid:
            self.nbr_state.zero_()
            self.rows_valid = True

    def drop_rows(self) -> None:
        """Positions were changed behind the rows' back (upload, direct writes to ``pos``): rebuild at the next launch."""
        self.rows_valid = False

    # -- kernels -----------------------------------------------------------------------------------
    def energy_full(self, pair_count: torch.Tensor | None = None) -> torch.Tensor:
        """Full potential energy of every replica; primes the EMT caches (``validate_simulation``).

        Replicas beyond the shared-memory capacity of the warp-per-replica kernels (> 1024 atoms: the Hamiltonian
        configurations) are evaluated by the neighbour-list kernels instead (energy only, no sweep caches)."""
        if self.n_max > 1024 and pair_count is None:
            self.forces_full()
            return self.energy
        with torch.cuda.device(self.device):
            b = self.struct()
            _lib.check(
                self.lib.qmc_energy_full(
                    C.byref(self.potential), C.byref(b), pair_count.data_ptr() if pair_count is not None else None, self.stream()
                )
            )
        return self.energy

    def atom_energies(self) -> torch.Tensor:
        """Per-atom energies [R][n_max] (``calc.results["energies"]`` of ASE's EMT / LennardJones) at the current positions; the
        energies and EMT caches are refreshed as by ``energy_full``."""
        out = torch.zeros((self.R, self.n_max), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            b = self.struct()
            _lib.check(self.lib.qmc_atom_energies(C.byref(self.potential), C.byref(b), out.data_ptr(), None, self.stream()))
        return out

    def forces_full(self) -> torch.Tensor:
        """Forces [R][3][n_max] (and energies) of every replica."""
        f = torch.zeros_like(self.pos)
        with torch.cuda.device(self.device):
            b = self.struct()
            ws = self.hmc_workspace(b)
            _lib.check(self.lib.qmc_forces_full(C.byref(self.potential), C.byref(b), f.data_ptr(), ws.data_ptr(), ws.numel(), self.stream()))
        self.raise_on_status()  # a truncated neighbour list must never pass as forces
        return f

    def hmc_workspace(self, b: _lib.Batch | None = None) -> torch.Tensor:
        """Device scratch of the Hamiltonian kernels (allocated once, reused)."""
        b = b if b is not None else self.struct()
        need = int(self.lib.qmc_hmc_workspace_bytes(C.byref(self.potential), C.byref(b)))
        if need < 0:
            _lib.check(need)
        ws = getattr(self, "_hmc_ws", None)
        if ws is None or ws.numel() < need:
            self._hmc_ws = ws = torch.zeros(max(need, 256), dtype=torch.uint8, device=self.device)
        return ws

    def raise_on_status(self) -> None:
        """Turn status bits the kernels recorded into the exceptions the reference would have raised."""
        st = int(torch.bitwise_or(self.status, 0).max().item()) if self.R else 0
        bits = 0
        if st:
            bits = int(np.bitwise_or.reduce(self.status.cpu().numpy()))
        if bits & _lib.STATUS_EXP_OVERFLOW:
            raise OverflowError("math range error")  # mc/criteria.py:108-110 -> math.exp
        if bits & _lib.STATUS_CAPACITY:
            raise ValueError("an insertion exceeded the atom capacity n_max of the batch")
        if bits & _lib.STATUS_CELL_TOO_SMALL:
            raise NotImplementedError("a periodic cell edge became shorter than the neighbour cutoff (plus the skin of the neighbour rows, if used)")
        if bits & _lib.STATUS_LIST_OVERFLOW:
            raise RuntimeError("neighbour list overflow")
        if bits & _lib.STATUS_ROW_SHIFT:
            raise RuntimeError("neighbour rows: interacting atoms whose unwrapped coordinates are more than 63 cells apart; "
                               "run with neighbour_rows=False")
