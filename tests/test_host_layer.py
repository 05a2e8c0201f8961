"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol, the quansino-facing
classes validate like the reference, and the GPU path refuses (never silently ignores) what it cannot run."""

import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    from quansino_b200 import _lib

    lib = _lib.load()
    header = (ROOT / "include" / "quansino_b200.h").read_text()
    declared = set(re.findall(r"\b(qmc_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.EXPORTS)
    for name in sorted(declared):
        assert getattr(lib, name) is not None, name
    assert lib.qmc_abi_version() == int(re.search(r"#define\s+QMC_ABI_VERSION\s+(\d+)", header).group(1)) == 9


def test_struct_layouts_match_the_header():
    """ctypes mirrors of the header structs: sizes as the C compiler lays them out."""
    import ctypes as C
    import subprocess
    import tempfile

    from quansino_b200 import _lib

    src = '#include <stdio.h>\n#include "quansino_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(qmc_potential_t), sizeof(qmc_move_t), sizeof(qmc_batch_t), sizeof(qmc_trace_t));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "t.c").write_text(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), "-o", f"{d}/t", f"{d}/t.c"], check=True)
        sizes = [int(x) for x in subprocess.run([f"{d}/t"], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_lib.Potential), C.sizeof(_lib.Move), C.sizeof(_lib.Batch), C.sizeof(_lib.Trace)]


def test_integration_stub_structs_match_the_library():
    """The ctypes structs printed in INTEGRATION.md are the ones a maintainer would paste: same sizes as the library's."""
    import ctypes as C

    from quansino_b200 import _lib

    text = (ROOT / "INTEGRATION.md").read_text()
    ns = {"C": C}
    for name in ("qmc_potential", "qmc_move", "qmc_batch"):
        m = re.search(rf"class {name}\(C\.Structure\):.*?(_fields_ = .*?)\n\n(?:class|def) ", text, re.S)
        assert m, name
        exec("class T(C.Structure):\n    " + m.group(1).replace("\n", "\n    "), ns)
        assert C.sizeof(ns["T"]) == C.sizeof({"qmc_potential": _lib.Potential, "qmc_move": _lib.Move, "qmc_batch": _lib.Batch}[name]), name


def test_no_gpu_means_loud_failure_not_fallback():
    import torch

    from quansino_b200 import EMT, BatchedAtoms, ReplicaBatch, _lib
    from quansino_b200.systems import fcc_cubic

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert _lib.load().qmc_device_count() == _lib.ERR_NO_DEVICE
    pos, cell = fcc_cubic("Cu", 3)
    atoms = BatchedAtoms.replicate(pos, cell, 2, calc=EMT())
    with pytest.raises(_lib.NoDeviceError):
        ReplicaBatch(atoms)


def test_product_never_imports_the_oracle():
    for path in (ROOT / "quansino_b200").rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{path} imports the oracle"
    for path in (ROOT / "quansino_b200" / "csrc").glob("*"):
        assert "qmc_oracle" not in path.read_text(), f"{path} references the oracle"


def test_emt_and_lj_descriptions_match_the_oracle_constants():
    from oracle.potentials import LJParams, emt_params
    from quansino_b200 import EMT, LennardJones

    for sym in ("Cu", "Pt", "Al", "Au"):
        p, ref = EMT().lower(sym), emt_params(sym)
        for a, b in [("E0", "E0"), ("s0", "s0"), ("V0", "V0"), ("eta2", "eta2"), ("kappa", "kappa"), ("lmbda", "lmbda"), ("rc", "rc"),
                     ("rc_list", "rc_list"), ("acut", "acut"), ("gamma1", "gamma1"), ("gamma2", "gamma2")]:  # fmt: skip
            assert getattr(p, a) == getattr(ref, b), (sym, a)
    p, ref = LennardJones(sigma=2.5, epsilon=0.1).lower("Pt"), LJParams(sigma=2.5, epsilon=0.1)
    assert (p.sigma, p.epsilon, p.lj_rc, p.lj_e0) == (ref.sigma, ref.epsilon, ref.cutoff, ref.e0)
    with pytest.raises(NotImplementedError):
        EMT().lower("Fe")
    with pytest.raises(NotImplementedError):
        EMT(asap_cutoff=True)


def test_moves_lower_like_the_reference_defaults():
    from quansino_b200 import _lib
    from quansino_b200.integrators import Verlet
    from quansino_b200.moves import CellMove, DisplacementMove, ExchangeMove, HamiltonianDisplacementMove
    from quansino_b200.operations import Ball, Box, IsotropicDeformation, Rotation, Sphere, Translation

    m = DisplacementMove(np.arange(10)).lower()
    assert (m.type, m.op, m.step) == (_lib.MOVE_DISPLACEMENT, _lib.OP_BALL, 0.1)  # default Ball(0.1), moves/displacement.py:214-224
    assert DisplacementMove(np.arange(4), Box(0.3)).lower().op == _lib.OP_BOX
    assert DisplacementMove(np.arange(4), Sphere(0.3)).lower().op == _lib.OP_SPHERE
    c = CellMove().lower()
    assert (c.type, c.op, c.step) == (_lib.MOVE_CELL, _lib.OP_ISOTROPIC, 0.05)  # moves/cell.py:112-122
    x = ExchangeMove(np.arange(4)).lower()
    assert (x.type, x.op, x.bias_insert) == (_lib.MOVE_EXCHANGE, _lib.OP_TRANSLATION, 0.5)
    h = HamiltonianDisplacementMove().lower()
    assert (h.type, h.op, h.n_steps) == (_lib.MOVE_HMC, _lib.OP_VERLET, 100) and abs(h.dt - 0.09822694788464063) < 1e-16
    assert HamiltonianDisplacementMove().max_attempts == 10 and DisplacementMove(np.arange(2)).max_attempts == 10000
    # refused, never ignored
    assert DisplacementMove(np.arange(4), Rotation()).lower().op == _lib.OP_ROTATION  # label groups rotate rigidly
    from quansino_b200.operations import TranslationRotation

    assert DisplacementMove(np.arange(4), Translation()).lower().op == _lib.OP_TRANSLATION  # the group's centroid is relocated
    assert DisplacementMove(np.arange(4), TranslationRotation()).lower().op == _lib.OP_TRANSLATION_ROTATION
    with pytest.raises(NotImplementedError):
        DisplacementMove(np.arange(4), Rotation() + Ball(0.1)).lower()
    with pytest.raises(NotImplementedError):
        CellMove(scale_atoms=False).lower()
    assert CellMove().lower().axes == 7
    assert CellMove(IsotropicDeformation(0.05, mask=np.diag([True, True, False]))).lower().axes == 3  # the slab case: z is fixed
    with pytest.raises(NotImplementedError):
        CellMove(IsotropicDeformation(0.05, mask=np.zeros((3, 3), dtype=bool))).lower()
    mv = DisplacementMove(np.arange(4))
    mv.check_move = lambda ctx: True
    with pytest.raises(NotImplementedError):
        mv.lower()
    with pytest.raises(NotImplementedError):
        Verlet(apply_constraints=False).lower()
    with pytest.raises(NotImplementedError):
        Ball(0.1).calculate(None)
    with pytest.raises(TypeError):
        DisplacementMove(np.arange(4)) + 3


def test_move_algebra_builds_composites_like_the_reference():
    """moves/core.py:185-243, moves/composite.py:62-151, tests/moves/test_displacement_moves.py of the reference."""
    from quansino_b200 import _lib
    from quansino_b200.moves import CellMove, CompositeDisplacementMove, CompositeMove, DisplacementMove
    from quansino_b200.operations import Ball, Box, Sphere

    lab = np.arange(6)
    a, b, c = DisplacementMove(lab, Ball(0.1)), DisplacementMove(lab, Box(0.2)), DisplacementMove(lab, Sphere(0.3))
    triple = a * 3
    assert type(triple) is CompositeDisplacementMove and len(triple) == 3 and triple[0] is a and list(triple) == [a, a, a]
    assert type(3 * a) is CompositeDisplacementMove
    mixed = a + b + c
    assert type(mixed) is CompositeDisplacementMove and mixed.moves == [a, b, c]
    assert type(mixed + a) is CompositeDisplacementMove and len(mixed + a) == 4
    assert type(a + (b + c)) is CompositeDisplacementMove and (a + (b + c)).moves == [a, b, c]
    assert len(mixed * 2) == 6 and type(mixed * 2) is CompositeDisplacementMove
    hybrid = a * 2 + CellMove()  # the docs' hybrid NPT move: a plain CompositeMove
    assert type(hybrid) is CompositeMove and len(hybrid) == 3
    assert type(a + CellMove()) is CompositeMove and type(CellMove() + a) is CompositeMove
    for bad in (0, -1, 1.5):
        with pytest.raises((ValueError, TypeError)):
            a * bad
    with pytest.raises(TypeError):
        mixed + 3
    # lowering: runs of identical sub-moves collapse into `repeat`, the entries are chained
    chain = (a * 3 + b + c * 2).lower_chain()
    assert [(m.op, m.step, m.repeat, m.chain) for m in chain] == [
        (_lib.OP_BALL, 0.1, 3, 1), (_lib.OP_BOX, 0.2, 1, 1), (_lib.OP_SPHERE, 0.3, 2, 0),
    ]  # fmt: skip
    assert [(m.type, m.op, m.repeat, m.chain) for m in hybrid.lower_chain()] == [
        (_lib.MOVE_DISPLACEMENT, _lib.OP_BALL, 2, 3), (_lib.MOVE_CELL, _lib.OP_ISOTROPIC, 1, 2),
    ]  # fmt: skip  (plain CompositeMove: chain bit 1; the CellMove closes the chain)
    assert [(m.repeat, m.chain) for m in CompositeMove([a, a]).lower_chain()] == [(2, 2)]
    with pytest.raises(NotImplementedError):
        (CellMove() + a).lower_chain()  # a cell move anywhere but last
    with pytest.raises(NotImplementedError):
        CompositeMove([a, CellMove(), CellMove()]).lower_chain()
    with pytest.raises(NotImplementedError):  # different label arrays inside one composite
        (a + DisplacementMove(np.array([0, 1, 2, -1, -1, -1]))).lower_chain()
    with pytest.raises(NotImplementedError):
        triple(None)
    d = mixed.to_dict()
    assert d["name"] == "CompositeDisplacementMove" and len(d["kwargs"]["moves"]) == 3


def test_label_validation():
    from quansino_b200.moves import DisplacementMove

    mv = DisplacementMove(np.array([0, -1, 1, 2, -1]))
    assert np.array_equal(mv.unique_labels, [0, 1, 2])
    lab = mv.validated_labels(2, 6, np.array([5, 5]))
    assert lab.shape == (2, 6) and np.array_equal(lab[0], [0, -1, 1, 2, -1, -1])
    assert not mv._grouped
    grp = DisplacementMove(np.array([40, 40, 7, -1, 300, 7]))  # label groups: rewritten as ranks of the sorted unique labels
    assert np.array_equal(grp.unique_labels, [7, 40, 300])
    assert np.array_equal(grp.validated_labels(1, 7, np.array([6]))[0], [1, 1, 0, -1, 2, 0, -1]) and grp._grouped
    with pytest.raises(ValueError):
        DisplacementMove(np.arange(3)).validated_labels(1, 8, np.array([5]))
    assert DisplacementMove(np.arange(5)).is_identity(np.array([5, 5]))
    assert not DisplacementMove(np.arange(5)).is_identity(np.array([5, 4]))


def test_cell_checks():
    from quansino_b200.batch import _check_cell

    assert _check_cell(np.diag([10.83, 10.83, 10.83])[None], (True,) * 3, 5.8768) is False
    with pytest.raises(NotImplementedError):
        _check_cell(np.diag([3.61, 3.61, 3.61])[None], (True,) * 3, 5.8768)  # the reference's 4-atom fixture
    assert _check_cell(np.diag([3.61, 3.61, 3.61])[None], (False,) * 3, 5.8768) is False
    # general cells: allowed (-> the general-cell kernels) as long as every periodic HEIGHT clears the cutoff
    tri = np.array([[10.0, 0, 0], [1.0, 10.0, 0], [0, 0, 10.0]])[None]
    assert _check_cell(tri, (True,) * 3, 5.0) is True
    flat = np.array([[10.0, 0, 0], [9.0, 5.2, 0], [0, 0, 10.0]])[None]  # edges of 10 A, but a height of 5.2 * 10 / |(9, 5.2)| = 5.0 A
    with pytest.raises(NotImplementedError, match="height"):
        _check_cell(flat, (True,) * 3, 5.1)
    assert _check_cell(flat, (False, False, True), 5.1) is True
    with pytest.raises(ValueError):
        _check_cell(np.zeros((1, 3, 3)), (True,) * 3, 5.0)


def test_shard_is_a_partition():
    from quansino_b200.mc.tempering import shard

    for n, w in [(64, 8), (4096, 8), (10, 4), (7, 8), (1, 1)]:
        blocks = [shard(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
        assert all(blocks[r][0] + blocks[r][1] == blocks[r + 1][0] for r in range(w - 1))
        assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def test_to_dict_from_dict_round_trips_through_the_registry():
    """``to_dict`` / ``from_dict`` with the class registry (src/quansino/registry.py, moves/core.py:151-183,
    moves/composite.py:218-247, utils/moves.py:62-102): the rebuilt objects lower to the same POD."""
    import json

    from quansino_b200.integrators import Verlet
    from quansino_b200.io.restart import _Encoder, _decode
    from quansino_b200.mc.criteria import CanonicalCriteria, IsobaricCriteria
    from quansino_b200.moves import CellMove, DisplacementMove, ExchangeMove, HamiltonianDisplacementMove
    from quansino_b200.operations import Ball, Box, IsotropicDeformation, Rotation, Sphere
    from quansino_b200.registry import get_class, get_typed_class, register_class
    from quansino_b200.utils.moves import MoveStorage

    def pod(m):
        vals = (getattr(m, f) for f, _ in m._fields_ if f != "labels")
        return tuple(tuple(v) if hasattr(v, "__len__") else v for v in vals)

    def through_json(d):
        return _decode(json.loads(json.dumps(d, cls=_Encoder)))

    lab = np.array([0, 1, -1, 2, 3])
    moves = [
        DisplacementMove(lab, Box(0.25)), DisplacementMove(lab, Sphere(0.05)), DisplacementMove(np.array([5, 5, 7, 7, -1]), Rotation()),
        CellMove(IsotropicDeformation(0.02, mask=np.diag([True, False, True]))), ExchangeMove(lab, bias_towards_insert=0.3),
        HamiltonianDisplacementMove(operation=Verlet(dt=2.0, max_steps=7)),
    ]  # fmt: skip
    for mv in moves:
        again = get_class(mv.to_dict()["name"]).from_dict(through_json(mv.to_dict()))
        assert type(again) is type(mv) and pod(again.lower()) == pod(mv.lower())
        if hasattr(mv, "labels"):
            assert np.array_equal(again.labels, mv.labels)
    combo = DisplacementMove(lab, Ball(0.1)) * 2 + DisplacementMove(lab, Box(0.2))
    again = get_class("CompositeDisplacementMove").from_dict(through_json(combo.to_dict()))
    assert [pod(m) for m in again.lower_chain()] == [pod(m) for m in combo.lower_chain()]
    hybrid = DisplacementMove(lab, Ball(0.1)) * 3 + CellMove()
    again = get_class("CompositeMove").from_dict(through_json(hybrid.to_dict()))
    assert [pod(m) for m in again.lower_chain()] == [pod(m) for m in hybrid.lower_chain()]
    st = MoveStorage(move=moves[3], criteria=IsobaricCriteria(), interval=2, probability=0.25, minimum_count=1)
    again = MoveStorage.from_dict(through_json(st.to_dict()))
    assert (again.interval, again.probability, again.minimum_count) == (2, 0.25, 1) and type(again.criteria) is IsobaricCriteria
    assert pod(again.move.lower()) == pod(moves[3].lower())
    assert get_typed_class("CanonicalCriteria", type(CanonicalCriteria()).__mro__[1]) is CanonicalCriteria
    with pytest.raises(KeyError):
        get_class("NoSuchMove")

    ops = Ball(0.1) + Box(0.2) * 2  # operations/core.py:72-116, operations/composite.py
    assert type(ops).__name__ == "CompositeOperation" and len(ops) == 3 and type(ops[1]) is Box and type(3 * Ball(0.1)).__name__ == "CompositeOperation"
    mv = DisplacementMove(lab, ops)
    low = mv.lower()
    assert (low.op, low.step, low.n_extra_ops, list(low.extra_op), list(low.extra_step)) == (0, 0.1, 2, [1, 1], [0.2, 0.2])
    again = get_class("DisplacementMove").from_dict(through_json(mv.to_dict()))
    assert pod(again.lower())[:4] == pod(low)[:4] and list(again.lower().extra_step) == [0.2, 0.2]
    with pytest.raises(NotImplementedError):
        DisplacementMove(lab, Ball(0.1) * 4).lower()
    with pytest.raises(NotImplementedError):
        DisplacementMove(lab, Ball(0.1) + Rotation()).lower()

    class MyBall(Ball):
        pass

    register_class(MyBall)
    assert get_class("MyBall") is MyBall


def _dummy_batch(lib_mod, with_n_exchange: bool):
    """A batch whose pointers are non-NULL but never dereferenced: qmc_sweep validates the table before it touches the device."""
    b = lib_mod.Batch()
    b.n_replicas, b.n_max = 2, 108
    for name in ("pos", "cell", "n_atoms", "energy", "temperature", "sigma1", "eatom", "counters", "status"):
        setattr(b, name, 4096)
    if with_n_exchange:
        b.n_exchange = 4096
    b.pbc = (1, 1, 1)
    return b


def test_sweep_refuses_tables_the_kernels_would_mishandle():
    """Advisor findings of round 1, as C-ABI behaviour: (1) label tables / composites without batch.n_exchange (the instantiation
    that serves them reads and writes it) are an invalid argument, not an illegal address; (2) the number of displacements of one
    composite attempt is bounded by the random-stream layout (round 1: 16, where the sub-move blocks ran into the extra draws)."""
    import ctypes as C

    from quansino_b200 import EMT, _lib

    lib = _lib.load()
    pot = EMT().lower("Cu")
    mv = (_lib.Move * 2)()
    mv[0].type, mv[0].op, mv[0].step, mv[0].probability = _lib.MOVE_DISPLACEMENT, _lib.OP_BALL, 0.1, 1.0
    mv[0].labels = 4096  # a label table
    rc = lib.qmc_sweep(C.byref(pot), C.byref(_dummy_batch(_lib, False)), mv, 1, 1, 0, 10, None, None)
    assert rc == _lib.ERR_INVALID and b"n_exchange" in lib.qmc_last_error()
    mv[0].labels = None
    # (2) sub-moves beyond the 16th have Philox blocks of their own (0x8000 ...) since the hybrid example of the reference needs
    # len(atoms) of them; the limit is the one the header states
    mv[0].repeat = _lib.QMC_MAX_COMPOSITE_DISPLACEMENTS + 1
    rc = lib.qmc_sweep(C.byref(pot), C.byref(_dummy_batch(_lib, True)), mv, 1, 1, 0, 10, None, None)
    assert rc == _lib.ERR_UNSUPPORTED and b"displace at most" in lib.qmc_last_error()
    mv[0].repeat, mv[0].chain = 1000, 3  # a chain counts as a whole
    mv[1].type, mv[1].op, mv[1].step, mv[1].repeat, mv[1].chain = _lib.MOVE_DISPLACEMENT, _lib.OP_BALL, 0.1, 25, 2
    rc = lib.qmc_sweep(C.byref(pot), C.byref(_dummy_batch(_lib, True)), mv, 2, 1, 0, 10, None, None)
    assert rc == _lib.ERR_UNSUPPORTED and b"displace at most" in lib.qmc_last_error()


def test_neighbour_row_geometry_is_exported():
    import ctypes as C

    from quansino_b200 import EMT, LennardJones, _lib

    lib = _lib.load()
    emt, lj = EMT().lower("Cu"), LennardJones(sigma=2.5, epsilon=0.1).lower("Pt")
    assert lib.qmc_nbr_row_words(C.byref(emt), 108) == 112 and lib.qmc_nbr_row_words(C.byref(emt), 500) == 160
    assert lib.qmc_nbr_row_words(C.byref(lj), 44) == 208
    assert lib.qmc_nbr_row_words(C.byref(emt), 2048) == _lib.ERR_UNSUPPORTED


def test_composite_exchange_move_algebra_and_lowering():
    """``ExchangeMove + ExchangeMove`` / ``* k`` build a ``CompositeExchangeMove`` (``moves/exchange.py:86,237-252``,
    ``tests/moves/test_composite_moves.py:264-296``); it lowers to ONE exchange entry with ``repeat`` = number of sub-moves."""
    from quansino_b200 import _lib, operations
    from quansino_b200.moves import CompositeExchangeMove, CompositeMove, DisplacementMove, ExchangeMove

    a, b = ExchangeMove(np.full(4, -1)), ExchangeMove(np.full(4, -1))
    comp = a + b
    assert isinstance(comp, CompositeExchangeMove) and len(comp) == 2 and comp.bias_towards_insert == 0.5
    four = comp * 2
    assert isinstance(four, CompositeExchangeMove) and len(four) == 4 and all(isinstance(m, ExchangeMove) for m in four)
    assert isinstance(a * 3, CompositeExchangeMove)
    assert type(a + DisplacementMove(np.arange(4))) is CompositeMove  # mixed kinds fall back to the plain composite
    with pytest.raises(ValueError):
        comp * 0
    with pytest.raises(TypeError):
        comp * 1.5
    four.bias_towards_insert = 0.25
    (entry,) = four.lower_chain()
    assert entry.type == _lib.MOVE_EXCHANGE and entry.op == _lib.OP_TRANSLATION and entry.repeat == 4 and entry.bias_insert == 0.25
    with pytest.raises(NotImplementedError, match="equal label arrays"):
        (a + ExchangeMove(np.array([-1, -1, 0, 1]))).lower_chain()
    with pytest.raises(NotImplementedError, match="Translation"):
        (a + ExchangeMove(np.full(4, -1), operations.TranslationRotation())).lower_chain()
    with pytest.raises(NotImplementedError, match="at most"):
        (a * (_lib.QMC_MAX_SUBMOVES + 1)).lower_chain()
    # serialisation through the registry keeps the composite's own bias
    from quansino_b200.registry import get_class

    d = four.to_dict()
    back = get_class(d["name"]).from_dict(d)
    assert isinstance(back, CompositeExchangeMove) and len(back) == 4 and back.bias_towards_insert == 0.25
    assert back.lower_chain()[0].repeat == 4


def test_generic_subscripts_and_protocols_like_the_reference():
    """The reference's drivers and moves are ``Generic`` (``Isobaric[Move, Criteria](...)`` in its examples,
    ``examples/orb-v3_isobaric_bulk_cu.py:23``) and ``quansino.protocols`` holds the structural types (``protocols.py:24-190``)."""
    from quansino_b200.integrators import Verlet
    from quansino_b200.mc import Canonical, ForceBias, GrandCanonical, HamiltonianCanonical, Isobaric, Isotension
    from quansino_b200.mc.criteria import IsobaricCriteria
    from quansino_b200.moves import CellMove, CompositeMove, DisplacementMove
    from quansino_b200.operations import Ball
    from quansino_b200.protocols import Criteria, Integrator, Move, Operation, Serializable
    from quansino_b200.utils.moves import MoveStorage

    for cls in (Canonical, ForceBias, GrandCanonical, HamiltonianCanonical, Isobaric, Isotension, CellMove, CompositeMove, MoveStorage):
        assert cls[Move, Criteria] is cls
    assert isinstance(DisplacementMove(np.arange(3)), Move) and isinstance(IsobaricCriteria(), Criteria)
    assert isinstance(Ball(0.1), Operation) and isinstance(Verlet(), Integrator) and isinstance(Ball(0.1), Serializable)
    assert not isinstance(object(), Move)


def test_search_molecules_and_import_paths_of_the_reference():
    """``utils/atoms.py:40-90`` (reference test ``tests/utils/test_atoms_utils.py``): bonded clusters share a label, sizes filter; and the
    module paths a script written for quansino imports exist here too."""
    import importlib
    from types import SimpleNamespace

    from quansino_b200.utils.atoms import search_molecules

    # three water-like trimers and two lone atoms in a periodic box; one trimer straddles the cell boundary
    mol = np.array([[0.0, 0.0, 0.0], [0.96, 0.0, 0.0], [-0.24, 0.93, 0.0]])
    pos = np.vstack([mol + [2.0, 2.0, 2.0], mol + [6.0, 6.0, 2.0], mol + [9.8, 1.0, 5.0], [[5.0, 1.0, 8.0]], [[1.0, 7.0, 7.0]]]) % 10.0  # wrapped
    atoms = SimpleNamespace(positions=pos, cell=np.eye(3) * 10.0, pbc=(True, True, True))
    lab = search_molecules(atoms, 1.2)
    assert list(lab) == [0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 4]
    assert list(search_molecules(atoms, 1.2, required_size=3)) == [0, 0, 0, 1, 1, 1, 2, 2, 2, -1, -1]
    assert list(search_molecules(atoms, 1.2, required_size=(1, 1))) == [-1] * 9 + [3, 4]
    open_box = SimpleNamespace(positions=pos, cell=np.eye(3) * 10.0, pbc=(False, False, False))
    assert len(set(search_molecules(open_box, 1.2))) == 6  # without periodicity the straddling trimer falls apart
    assert list(search_molecules(atoms, [0.6] * 11)) == [0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 4]  # per-atom radii
    for path, names in {
        "quansino_b200.protocols": ["Move", "Criteria", "Operation", "Integrator"], "quansino_b200.type_hints": ["IntegerArray", "Stress", "Deformation"],
        "quansino_b200.mc.driver": ["Driver", "SingleDriver", "MultiDriver"], "quansino_b200.io.file": ["ObserverManager"],
        "quansino_b200.integrators.core": ["BaseIntegrator"], "quansino_b200.utils.strings": ["get_auto_header_format"],
        "quansino_b200.moves.displacement": ["CompositeDisplacementMove", "DisplacementMove", "HamiltonianDisplacementMove"],
        "quansino_b200.moves.exchange": ["CompositeExchangeMove", "ExchangeMove"], "quansino_b200.mc.isotension": ["Isotension"],
        "quansino_b200.mc.contexts": ["HamiltonianContext", "HamiltonianDeformationContext", "HamiltonianExchangeContext", "MultiContexts"],
        "quansino_b200.registry": ["get_class_name", "get_class", "register_class"],
    }.items():  # fmt: skip
        module = importlib.import_module(path)
        for name in names:
            assert getattr(module, name) is not None, (path, name)
