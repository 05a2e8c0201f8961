"""ctypes binding of ``oracle/qmc_oracle.c``.  TEST INFRASTRUCTURE (see oracle/__init__.py)."""

from __future__ import annotations

import ctypes as C
import subprocess
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from oracle.potentials import EMT_BETA, EMTParams, LJParams, emt_params

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libqmc_oracle.so"

POT_EMT, POT_LJ = 0, 1
MOVE_DISPLACEMENT, MOVE_CELL, MOVE_EXCHANGE, MOVE_HMC = 0, 1, 2, 3
OP_BALL, OP_BOX, OP_SPHERE, OP_TRANSLATION, OP_ISOTROPIC, OP_VERLET, OP_ROTATION, OP_TRANSLATION_ROTATION = 0, 1, 2, 3, 4, 5, 6, 7
OP_ANISOTROPIC, OP_SHAPE = 8, 9
LABEL_NONE = -(2**31)


class Potential(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32)] + [
        (name, C.c_double)
        for name in (
            "E0", "s0", "V0", "eta2", "kappa", "lmbda", "beta", "rc", "rc_list", "acut", "gamma1", "gamma2",
            "sigma", "epsilon", "lj_rc", "lj_e0",
        )
    ]  # fmt: skip


class Move(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("op", C.c_int32),
        ("step", C.c_double),
        ("probability", C.c_double),
        ("bias_insert", C.c_double),
        ("dt", C.c_double),
        ("n_steps", C.c_int32),
        ("default_label", C.c_int32),
        ("repeat", C.c_int32),
        ("chain", C.c_int32),
        ("axes", C.c_int32),
        ("minimum_count", C.c_int32),
        ("max_cycles", C.c_int32),
        ("n_extra_ops", C.c_int32),
        ("extra_op", C.c_int32 * 2),
        ("_pad", C.c_int32),
        ("extra_step", C.c_double * 2),
        ("labels", C.POINTER(C.c_int32)),
    ]


class State(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int32),
        ("n_max", C.c_int32),
        ("pos", C.POINTER(C.c_double)),
        ("momenta", C.POINTER(C.c_double)),
        ("cell", C.c_double * 9),
        ("pbc", C.c_int32 * 3),
        ("mass", C.c_double),
        ("temperature", C.c_double),
        ("pressure", C.c_double),
        ("chemical_potential", C.c_double),
        ("accessible_volume", C.c_double),
        ("exchange_mass", C.c_double),
        ("exchange_pos", C.c_double * 3),
        ("n_exchange", C.c_int32),
        ("_pad", C.c_int32),
        ("last_energy", C.c_double),
        ("last_kinetic", C.c_double),
        ("n_exchange_atoms", C.c_int32),
        ("_pad2", C.c_int32),
        ("exchange_mol", C.c_double * 24),
        ("isotension", C.c_int32),
        ("_pad3", C.c_int32),
        ("external_stress", C.c_double * 9),
    ]


class Trace(C.Structure):
    _fields_ = [
        ("move", C.POINTER(C.c_int8)),
        ("accepted", C.POINTER(C.c_int8)),
        ("energy", C.POINTER(C.c_double)),
        ("delta", C.POINTER(C.c_double)),
        ("n_atoms", C.POINTER(C.c_int32)),
        ("moved", C.POINTER(C.c_int32)),
    ]


def build(force: bool = False) -> Path:
    """Compile the C restatement with the committed Makefile (gcc only, no reference sources involved)."""
    src = _HERE / "qmc_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < max(src.stat().st_mtime, (_HERE / "qmc_oracle.h").stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        L.qo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.qo_philox4x32_10.restype = None
        L.qo_uniform_pair.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
        L.qo_uniform_pair.restype = None
        L.qo_normal.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.qo_normal.restype = C.c_double
        L.qo_energy.argtypes = [C.POINTER(Potential), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
        L.qo_energy.restype = C.c_int64
        L.qo_mc_run.argtypes = [C.POINTER(Potential), C.POINTER(State), C.POINTER(Move), C.c_int32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int32, C.POINTER(Trace), C.POINTER(C.c_int64)]
        L.qo_mc_run.restype = C.c_int
        L.qo_debroglie_wavelength.argtypes = [C.c_double, C.c_double]
        L.qo_debroglie_wavelength.restype = C.c_double
        L.qo_gc_acceptance.argtypes = [C.c_double, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double]
        L.qo_gc_acceptance.restype = C.c_double
        L.qo_expm_sym3.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.qo_expm_sym3.restype = None
        L.qo_pt_swap_accept.argtypes = [C.c_double] * 5
        L.qo_pt_swap_accept.restype = C.c_int
        _lib = L
    return _lib


def make_potential(par) -> Potential:
    p = Potential()
    if isinstance(par, EMTParams):
        p.kind = POT_EMT
        p.E0, p.s0, p.V0, p.eta2, p.kappa, p.lmbda = par.E0, par.s0, par.V0, par.eta2, par.kappa, par.lmbda
        p.beta, p.rc, p.rc_list, p.acut, p.gamma1, p.gamma2 = EMT_BETA, par.rc, par.rc_list, par.acut, par.gamma1, par.gamma2
    elif isinstance(par, LJParams):
        p.kind = POT_LJ
        p.sigma, p.epsilon, p.lj_rc, p.lj_e0 = par.sigma, par.epsilon, par.cutoff, par.e0
    else:
        raise TypeError(par)
    return p


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().qo_philox4x32_10(c, k, o)
    return tuple(o)


def uniform_pair(seed, attempt, replica, block):
    o = (C.c_double * 2)()
    lib().qo_uniform_pair(seed, attempt, replica, block, o)
    return o[0], o[1]


def energy(par, positions, cell, pbc, want_forces=False, want_sigma=False):
    pos = np.ascontiguousarray(positions, dtype=np.float64)
    n = len(pos)
    cl = np.ascontiguousarray(np.asarray(cell, dtype=np.float64).reshape(9))
    pb = np.ascontiguousarray(np.asarray(pbc, dtype=np.int32))
    e = C.c_double()
    f = np.zeros((n, 3)) if want_forces else None
    s = np.zeros(n) if want_sigma else None
    pot = make_potential(par)
    npairs = lib().qo_energy(
        C.byref(pot), n, pos.ctypes.data, cl.ctypes.data, pb.ctypes.data, C.byref(e),
        f.ctypes.data if f is not None else None, s.ctypes.data if s is not None else None,
    )  # fmt: skip
    if npairs < 0:
        raise RuntimeError("qo_energy failed")
    return {"energy": e.value, "forces": f, "sigma1": s, "n_pairs": int(npairs)}


@dataclass
class MoveSpec:
    type: int
    op: int
    step: float = 0.0
    probability: float = 1.0
    bias_insert: float = 0.5
    dt: float = 0.0
    n_steps: int = 0
    default_label: int = LABEL_NONE
    labels: np.ndarray | None = None  # int32 [n_max]
    repeat: int = 1  # CompositeDisplacementMove: consecutive copies of this sub-move
    chain: int = 0  # bit 0: the next MoveSpec belongs to the same composite move; bit 1: plain CompositeMove
    axes: int = 0  # CellMove: deformed axes (bit mask), 0 = all
    minimum_count: int = 0  # forced occurrences per step
    max_cycles: int = 0  # cycles per step (needed with forced moves)
    extra_ops: tuple = ()  # CompositeOperation: up to two further (op, step) pairs


@dataclass
class Replica:
    """One replica's Atoms + Context fields, owned as numpy arrays and lent to the C oracle."""

    positions: np.ndarray  # (n_max, 3), first n_atoms rows valid
    cell: np.ndarray
    pbc: tuple = (True, True, True)
    n_atoms: int | None = None
    mass: float = 63.546
    temperature: float = 300.0
    pressure: float = 0.0
    chemical_potential: float = 0.0
    accessible_volume: float | None = None
    exchange_mass: float = 0.0
    exchange_pos: tuple = (0.0, 0.0, 0.0)
    exchange_mol: np.ndarray | None = None  # (k, 3) positions of a molecular exchange species (k <= 8); None: the single atom
    n_exchange: int = 0
    external_stress: np.ndarray | None = None  # (3, 3) eV / A^3: cell moves are judged by IsotensionCriteria; None: IsobaricCriteria
    last_energy: float = float("nan")
    last_kinetic: float = float("nan")
    momenta: np.ndarray | None = None
    counters: np.ndarray = field(default_factory=lambda: np.zeros(4, dtype=np.int64))

    def __post_init__(self):
        self.positions = np.ascontiguousarray(self.positions, dtype=np.float64)
        self.cell = np.ascontiguousarray(np.asarray(self.cell, dtype=np.float64).reshape(3, 3))
        if self.n_atoms is None:
            self.n_atoms = len(self.positions)
        if self.accessible_volume is None:
            self.accessible_volume = abs(float(np.linalg.det(self.cell)))


def expm_sym3(a: np.ndarray) -> np.ndarray:
    """exp of a symmetric 3 x 3 matrix by the oracle's Jacobi rotations."""
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(9)
    out = np.zeros(9)
    lib().qo_expm_sym3(a.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out.reshape(3, 3)


def mc_run(par, rep: Replica, moves: list[MoveSpec], seed: int, replica_id: int, attempt_base: int, n_attempts: int):
    """Run ``n_attempts`` cycles on ``rep`` in place; returns the per-attempt trace as a dict of arrays."""
    n_max = len(rep.positions)
    st = State()
    st.n_atoms, st.n_max = rep.n_atoms, n_max
    st.pos = rep.positions.ctypes.data_as(C.POINTER(C.c_double))
    if rep.momenta is not None:
        rep.momenta = np.ascontiguousarray(rep.momenta, dtype=np.float64)
        st.momenta = rep.momenta.ctypes.data_as(C.POINTER(C.c_double))
    st.cell = (C.c_double * 9)(*rep.cell.reshape(9))
    st.pbc = (C.c_int32 * 3)(*[int(b) for b in rep.pbc])
    st.mass, st.temperature, st.pressure = rep.mass, rep.temperature, rep.pressure
    st.chemical_potential, st.accessible_volume = rep.chemical_potential, rep.accessible_volume
    st.exchange_mass = rep.exchange_mass
    st.exchange_pos = (C.c_double * 3)(*rep.exchange_pos)
    if rep.exchange_mol is not None:
        mol = np.asarray(rep.exchange_mol, dtype=np.float64).reshape(-1, 3)
        if len(mol) > 8:
            raise ValueError("the oracle holds exchange molecules of at most 8 atoms")
        st.n_exchange_atoms = len(mol)
        for q, v in enumerate(mol.reshape(-1)):
            st.exchange_mol[q] = float(v)
    st.n_exchange = rep.n_exchange
    if rep.external_stress is not None:
        st.isotension = 1
        st.external_stress = (C.c_double * 9)(*np.asarray(rep.external_stress, dtype=np.float64).reshape(9))
    st.last_energy, st.last_kinetic = rep.last_energy, rep.last_kinetic

    mv = (Move * len(moves))()
    for i, m in enumerate(moves):
        mv[i].type, mv[i].op, mv[i].step, mv[i].probability = m.type, m.op, m.step, m.probability
        mv[i].bias_insert, mv[i].dt, mv[i].n_steps, mv[i].default_label = m.bias_insert, m.dt, m.n_steps, m.default_label
        mv[i].repeat, mv[i].chain, mv[i].axes = m.repeat, m.chain, m.axes
        mv[i].minimum_count, mv[i].max_cycles = m.minimum_count, m.max_cycles
        mv[i].n_extra_ops = len(m.extra_ops)
        for q, (op, step) in enumerate(m.extra_ops):
            mv[i].extra_op[q], mv[i].extra_step[q] = op, step
        if m.labels is not None:
            assert m.labels.dtype == np.int32 and len(m.labels) >= n_max and m.labels.flags.c_contiguous
            mv[i].labels = m.labels.ctypes.data_as(C.POINTER(C.c_int32))

    tr_move = np.zeros(n_attempts, dtype=np.int8)
    tr_acc = np.zeros(n_attempts, dtype=np.int8)
    tr_e = np.zeros(n_attempts)
    tr_d = np.zeros(n_attempts)
    tr_n = np.zeros(n_attempts, dtype=np.int32)
    tr_m = np.zeros(n_attempts, dtype=np.int32)
    tr = Trace()
    tr.move = tr_move.ctypes.data_as(C.POINTER(C.c_int8))
    tr.accepted = tr_acc.ctypes.data_as(C.POINTER(C.c_int8))
    tr.energy = tr_e.ctypes.data_as(C.POINTER(C.c_double))
    tr.delta = tr_d.ctypes.data_as(C.POINTER(C.c_double))
    tr.n_atoms = tr_n.ctypes.data_as(C.POINTER(C.c_int32))
    tr.moved = tr_m.ctypes.data_as(C.POINTER(C.c_int32))
    pot = make_potential(par)
    rc = lib().qo_mc_run(
        C.byref(pot), C.byref(st), mv, len(moves), seed, replica_id, attempt_base, n_attempts, C.byref(tr),
        rep.counters.ctypes.data_as(C.POINTER(C.c_int64)),
    )  # fmt: skip
    if rc != 0:
        raise RuntimeError(f"qo_mc_run failed with code {rc}")
    rep.n_atoms = st.n_atoms
    rep.cell = np.array(st.cell[:], dtype=np.float64).reshape(3, 3)
    rep.n_exchange = st.n_exchange
    rep.last_energy, rep.last_kinetic = st.last_energy, st.last_kinetic
    return {"move": tr_move, "accepted": tr_acc, "energy": tr_e, "delta": tr_d, "n_atoms": tr_n, "moved": tr_m}


def fbmc_step(par, positions, cell, pbc, mass, temperature, delta, seed, replica_id, step):
    """One ForceBias step in place on ``positions`` (n, 3); returns (energy after, max |gamma|, rejection iterations)."""
    L = lib()
    L.qo_fbmc_step.restype = C.c_int
    pos = np.ascontiguousarray(positions, dtype=np.float64)
    assert pos is positions or np.shares_memory(pos, positions)
    pot = make_potential(par)
    cellf = np.ascontiguousarray(np.asarray(cell, dtype=np.float64).reshape(9))
    pb = (C.c_int32 * 3)(*[int(b) for b in pbc])
    e, g, it = C.c_double(), C.c_double(), C.c_int32()
    rc = L.qo_fbmc_step(C.byref(pot), C.c_int32(len(pos)), pos.ctypes.data_as(C.POINTER(C.c_double)), cellf.ctypes.data_as(C.POINTER(C.c_double)),
                        pb, C.c_double(mass), C.c_double(temperature), C.c_double(delta), C.c_uint64(seed), C.c_uint32(replica_id),
                        C.c_uint64(step), C.byref(e), C.byref(g), C.byref(it))  # fmt: skip
    if rc != 0:
        raise RuntimeError(f"qo_fbmc_step failed with code {rc}")
    return e.value, g.value, it.value


__all__ = [
    "LABEL_NONE", "MOVE_CELL", "MOVE_DISPLACEMENT", "MOVE_EXCHANGE", "MOVE_HMC", "OP_BALL", "OP_BOX", "OP_ISOTROPIC",
    "OP_ANISOTROPIC", "OP_ROTATION", "OP_SHAPE", "OP_SPHERE", "OP_TRANSLATION", "OP_TRANSLATION_ROTATION", "OP_VERLET", "MoveSpec", "Replica", "build", "emt_params", "energy", "expm_sym3", "fbmc_step", "lib", "mc_run",
    "philox", "uniform_pair",
]  # fmt: skip
