"""ctypes binding of ``libquansino_b200.so`` (the C ABI declared in ``include/quansino_b200.h``).

There is no CPU fallback: if the library is missing, or a compute entry point finds no CUDA device, an
exception is raised.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(__import__("os").environ.get("QUANSINO_B200_LIB", _PKG / "libquansino_b200.so"))

QMC_MAX_MOVES = 4
QMC_MAX_SUBMOVES = 16
QMC_MAX_COMPOSITE_DISPLACEMENTS = 1024
LABEL_NONE = -(2**31)

POT_EMT, POT_LJ = 0, 1
MOVE_DISPLACEMENT, MOVE_CELL, MOVE_EXCHANGE, MOVE_HMC = 0, 1, 2, 3
(FIELD_POS, FIELD_CELL, FIELD_N_ATOMS, FIELD_ENERGY, FIELD_TEMPERATURE, FIELD_SIGMA1, FIELD_EATOM, FIELD_MOMENTA, FIELD_KINETIC,
 FIELD_N_EXCHANGE, FIELD_COUNTERS, FIELD_STATUS, FIELD_PAIR_EVALS, FIELD_LABELS0) = range(14)
OP_BALL, OP_BOX, OP_SPHERE, OP_TRANSLATION, OP_ISOTROPIC, OP_VERLET, OP_ROTATION, OP_TRANSLATION_ROTATION = 0, 1, 2, 3, 4, 5, 6, 7
OP_ANISOTROPIC, OP_SHAPE = 8, 9

NBR_SKIN_DEFAULT = 0.7  # include/quansino_b200.h: QMC_NBR_SKIN_DEFAULT
ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE = -1, -2, -3, -4
STATUS_EXP_OVERFLOW, STATUS_CAPACITY, STATUS_CELL_TOO_SMALL, STATUS_LIST_OVERFLOW, STATUS_ROW_SHIFT = 1, 2, 4, 8, 16

EXPORTS = (
    "qmc_abi_version", "qmc_last_error", "qmc_device_count", "qmc_sweep_smem_per_replica", "qmc_energy_full", "qmc_atom_energies",
    "qmc_forces_full", "qmc_sweep", "qmc_hmc", "qmc_hmc_workspace_bytes", "qmc_pt_swap", "qmc_fp64_peak",
    "qmc_sweep_host", "qmc_debug_math", "qmc_fbmc_step", "qmc_fbmc_step_v", "qmc_batch_create", "qmc_batch_destroy", "qmc_batch_upload",
    "qmc_batch_download", "qmc_batch_view", "qmc_batch_labels", "qmc_nbr_row_words", "qmc_batch_enable_rows", "qmc_pt_exchange", "qmc_reduce_observables",
)  # fmt: skip


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
        ("labels", C.c_void_p),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("n_replicas", C.c_int32),
        ("n_max", C.c_int32),
        ("replica_offset", C.c_int32),
        ("pbc", C.c_int32 * 3),
        ("cells_uniform", C.c_int32),
        ("nbr_row_cap", C.c_int32),
        ("cell_diag", C.c_double * 3),
        ("pos", C.c_void_p),
        ("cell", C.c_void_p),
        ("n_atoms", C.c_void_p),
        ("energy", C.c_void_p),
        ("temperature", C.c_void_p),
        ("sigma1", C.c_void_p),
        ("eatom", C.c_void_p),
        ("momenta", C.c_void_p),
        ("kinetic", C.c_void_p),
        ("n_exchange", C.c_void_p),
        ("counters", C.c_void_p),
        ("status", C.c_void_p),
        ("pair_evals", C.c_void_p),
        ("nbr_rows", C.c_void_p),
        ("nbr_ref", C.c_void_p),
        ("nbr_state", C.c_void_p),
        ("nbr_skin", C.c_double),
        ("mass", C.c_double),
        ("pressure", C.c_double),
        ("chemical_potential", C.c_double),
        ("accessible_volume", C.c_double),
        ("exchange_mass", C.c_double),
        ("exchange_pos", C.c_double * 3),
        ("n_exchange_atoms", C.c_int32),
        ("_pad2", C.c_int32),
        ("exchange_mol", C.c_double * 24),
        ("triclinic", C.c_int32),
        ("isotension", C.c_int32),
        ("external_stress", C.c_double * 9),
    ]


class Trace(C.Structure):
    _fields_ = [
        ("move", C.c_void_p),
        ("accepted", C.c_void_p),
        ("energy", C.c_void_p),
        ("delta", C.c_void_p),
        ("moved", C.c_void_p),
    ]


class NoDeviceError(RuntimeError):
    """Raised when a compute entry point is called without a CUDA device (there is no CPU fallback)."""


_lib = None


def load() -> C.CDLL:
    """Load the C-ABI library built by ``python -m quansino_b200.build``; fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m quansino_b200.build` "
            "(quansino_b200 has no CPU fallback and will not run without its CUDA library)"
        )
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    L.qmc_abi_version.restype = C.c_int
    L.qmc_last_error.restype = C.c_char_p
    L.qmc_device_count.restype = C.c_int
    L.qmc_sweep_smem_per_replica.argtypes = [C.POINTER(Potential), i32]
    L.qmc_sweep_smem_per_replica.restype = i64
    L.qmc_energy_full.argtypes = [C.POINTER(Potential), C.POINTER(Batch), vp, vp]
    L.qmc_atom_energies.argtypes = [C.POINTER(Potential), C.POINTER(Batch), vp, vp, vp]
    L.qmc_atom_energies.restype = C.c_int
    L.qmc_forces_full.argtypes = [C.POINTER(Potential), C.POINTER(Batch), vp, vp, i64, vp]
    L.qmc_sweep.argtypes = [C.POINTER(Potential), C.POINTER(Batch), C.POINTER(Move), i32, u64, u64, i32, C.POINTER(Trace), vp]
    L.qmc_hmc.argtypes = [C.POINTER(Potential), C.POINTER(Batch), C.POINTER(Move), u64, u64, i32, C.POINTER(Trace), vp, i64, vp]
    L.qmc_hmc_workspace_bytes.argtypes = [C.POINTER(Potential), C.POINTER(Batch)]
    L.qmc_hmc_workspace_bytes.restype = i64
    L.qmc_pt_swap.argtypes = [vp, vp, i32, u64, u64, vp, vp]
    L.qmc_pt_exchange.argtypes = [vp, vp, vp, i32, u64, u64, vp, vp, vp]
    L.qmc_pt_exchange.restype = C.c_int
    L.qmc_reduce_observables.argtypes = [vp, vp, i32, vp]
    L.qmc_reduce_observables.restype = C.c_int
    L.qmc_fp64_peak.argtypes = [C.POINTER(C.c_double), vp]
    L.qmc_sweep_host.argtypes = [C.POINTER(Potential), C.POINTER(Batch), C.POINTER(Move), i32, u64, u64, i32, C.POINTER(Trace)]
    L.qmc_fbmc_step.argtypes = [C.POINTER(Potential), C.POINTER(Batch), C.c_double, u64, u64, vp, i32, vp, vp, vp, i64, vp]
    L.qmc_fbmc_step.restype = C.c_int
    L.qmc_fbmc_step_v.argtypes = [C.POINTER(Potential), C.POINTER(Batch), C.c_double, vp, u64, u64, vp, i32, vp, vp, vp, i64, vp]
    L.qmc_fbmc_step_v.restype = C.c_int
    L.qmc_batch_create.argtypes = [i32, i32, i32, i32, C.POINTER(vp)]
    L.qmc_batch_destroy.argtypes = [vp]
    L.qmc_batch_upload.argtypes = [vp, i32, vp, i64]
    L.qmc_batch_download.argtypes = [vp, i32, vp, i64]
    L.qmc_batch_view.argtypes = [vp, C.POINTER(Batch)]
    L.qmc_batch_enable_rows.argtypes = [vp, C.POINTER(Potential)]
    L.qmc_batch_enable_rows.restype = C.c_int
    L.qmc_batch_labels.argtypes = [vp, i32]
    L.qmc_batch_labels.restype = vp
    for name in ("qmc_batch_create", "qmc_batch_destroy", "qmc_batch_upload", "qmc_batch_download", "qmc_batch_view"):
        getattr(L, name).restype = C.c_int
    L.qmc_nbr_row_words.argtypes = [C.POINTER(Potential), i32]
    L.qmc_nbr_row_words.restype = i32
    L.qmc_debug_math.argtypes = [i32, vp, vp, i32, vp]
    L.qmc_debug_math.restype = C.c_int
    for name in ("qmc_energy_full", "qmc_forces_full", "qmc_sweep", "qmc_hmc", "qmc_pt_swap", "qmc_fp64_peak", "qmc_sweep_host"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    """Map a C status code onto the exception type the reference would raise."""
    if rc >= 0:
        return
    msg = load().qmc_last_error().decode(errors="replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_NO_DEVICE:
        raise NoDeviceError(msg + " (quansino_b200 has no CPU fallback)")
    raise RuntimeError(msg)
