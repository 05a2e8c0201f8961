"""Randomised differential test: seeded random systems and move tables through the CUDA path and the C oracle.

Every case draws a potential (EMT / Lennard-Jones), a cell (cubic, slab, sheared, cluster in a box), ragged replica sizes, and a
move table out of the kinds the reference offers -- displacement (Ball / Box / Sphere, optionally a CompositeOperation or a
`move * k` composite), cell moves (the three deformation operations, Isobaric or Isotension criteria), plain exchange moves -- with
random step sizes and probabilities, runs it for a few hundred attempts over two launches, and compares every attempt of every
replica.  The point is breadth: combinations nobody wrote a dedicated test for."""

import os

import numpy as np
import pytest

from oracle import capi
from oracle.potentials import LJParams, emt_params
from quansino_b200.systems import fcc_cubic, rattle

pytestmark = pytest.mark.gpu

# QMC_FUZZ_SCALE=4 runs four times as many seeds per family (a one-off hunt; the default keeps the suite short)
SCALE = max(1, int(os.environ.get("QMC_FUZZ_SCALE", "1")))
OFFSET = int(os.environ.get("QMC_FUZZ_OFFSET", "0"))  # first seed of a hunt (0: the committed seeds)
BAR = 6.241509125883258e-07
OPS = {"Ball": capi.OP_BALL, "Box": capi.OP_BOX, "Sphere": capi.OP_SPHERE}
CELL_OPS = {"Isotropic": capi.OP_ISOTROPIC, "Anisotropic": capi.OP_ANISOTROPIC, "Shape": capi.OP_SHAPE}


def _case(seed):
    """A reproducible random case description (plain dict)."""
    rng = np.random.default_rng(10_000 + seed)
    c = {"seed": seed, "pot": rng.choice(["emt", "emt", "lj"]), "repeat": int(rng.choice([3, 3, 4])), "R": 3}
    c["cell_kind"] = str(rng.choice(["cubic", "cubic", "slab", "sheared", "sheared"]))
    c["temperature"] = float(rng.choice([300.0, 900.0, 2500.0]))
    kinds = ["disp"]
    periodic = c["cell_kind"] in ("cubic", "sheared")
    if rng.random() < 0.35:
        kinds.append("exchange")  # (the reference's GrandCanonical has no pressure: exchange and cell moves do not meet)
    elif periodic and rng.random() < 0.7:
        kinds.append("cell")
    c["kinds"] = kinds
    c["disp_op"] = str(rng.choice(["Ball", "Box", "Sphere"]))
    c["disp_step"] = float(rng.uniform(0.03, 0.2))
    # extras of the displacement move live in the diagonal-cell kernels only, and not next to exchange moves
    plain = c["cell_kind"] != "sheared" and "exchange" not in kinds
    c["composite_k"] = int(rng.choice([1, 1, 2, 3])) if plain and "cell" not in kinds else 1
    c["extra_op"] = (str(rng.choice(["Ball", "Box", "Sphere"])), float(rng.uniform(0.02, 0.1))) if plain and c["composite_k"] == 1 and rng.random() < 0.4 else None
    c["cell_op"] = str(rng.choice(["Isotropic", "Anisotropic", "Shape"])) if "cell" in kinds else None
    c["cell_max"] = float(rng.uniform(0.003, 0.012))
    c["iso_mask"] = [bool(b) for b in rng.random(3) < 0.75] if c["cell_op"] == "Isotropic" else [True] * 3
    if not any(c["iso_mask"]):
        c["iso_mask"][int(rng.integers(3))] = True
    c["isotension"] = bool("cell" in kinds and rng.random() < 0.5)
    c["stress"] = (rng.uniform(-2e-3, 2e-3, (3, 3)) * 0.5).tolist()
    c["pressure"] = float(rng.uniform(0.0, 2000.0) * BAR)
    p = rng.uniform(0.2, 1.0, len(kinds))
    c["probabilities"] = (p / p.sum()).tolist()
    c["mu"] = float(rng.uniform(-4.2, -3.5) if c["pot"] == "emt" else rng.uniform(-4.8, -3.6))  # (where deletions start to be accepted)
    c["ragged"] = bool("exchange" not in kinds and rng.random() < 0.5)
    c["n_attempts"] = 90
    return c


def _system(c):
    rng = np.random.default_rng(20_000 + c["seed"])
    pos0, cell0 = fcc_cubic("Cu", c["repeat"])
    n = len(pos0)
    R = c["R"]
    n_max = n + (16 if "exchange" in c["kinds"] else 0)
    pos, cells = np.zeros((R, n_max, 3)), np.zeros((R, 3, 3))
    pbc = (True, True, True)
    for r in range(R):
        cell = cell0.copy()
        p = rattle(pos0, 0.05, 3000 + 10 * c["seed"] + r)
        if c["cell_kind"] == "sheared":
            S = np.eye(3) + rng.uniform(-0.08, 0.08, (3, 3)) * (1 - np.eye(3))
            frac = np.linalg.solve(cell0.T, p.T).T
            cell = S @ cell0
            p = frac @ cell
        elif c["cell_kind"] == "slab":
            cell[2, 2] *= 1.8
            pbc = (True, True, False)
        pos[r, :n], cells[r] = p, cell
    counts = np.full(R, n, dtype=np.int32)
    if c["ragged"]:
        counts = np.array([n, n - int(rng.integers(5, 40)), n - int(rng.integers(1, 5))], dtype=np.int32)
        for r in range(R):
            pos[r, counts[r] :] = 0.0
    return pos, cells, counts, pbc, n, n_max


def _run_gpu(c, pos, cells, counts, pbc, n, n_max):
    from quansino_b200 import EMT, BatchedAtoms, LennardJones, operations
    from quansino_b200.mc import Canonical, GrandCanonical, Isobaric, Isotension
    from quansino_b200.mc.criteria import CanonicalCriteria
    from quansino_b200.moves import CellMove, DisplacementMove, ExchangeMove

    calc = EMT() if c["pot"] == "emt" else LennardJones(sigma=2.3, epsilon=0.2, rc=6.0)
    atoms = BatchedAtoms(pos.copy(), cells, counts, "Cu", pbc, calc)
    op = getattr(operations, c["disp_op"])(c["disp_step"])
    if c["extra_op"]:
        op = op + getattr(operations, c["extra_op"][0])(c["extra_op"][1])
    disp = DisplacementMove(np.arange(n_max if "exchange" not in c["kinds"] else n), op)
    common = dict(temperature=c["temperature"], seed=4242 + c["seed"], trace=True, resync_interval=0, max_cycles=c["n_attempts"])
    cell_move = None
    if "cell" in c["kinds"]:
        name = f"{c['cell_op']}Deformation"
        kw = {"mask": np.diag(c["iso_mask"])} if c["cell_op"] == "Isotropic" else {}
        cell_move = CellMove(getattr(operations, name)(c["cell_max"], **kw))
    if "exchange" in c["kinds"]:
        mc = GrandCanonical(atoms, exchange_atoms=("Cu", (0.0, 0.0, 0.0)), chemical_potential=c["mu"], number_of_exchange_particles=n,
                            default_displacement_move=disp, default_exchange_move=ExchangeMove(np.arange(n)), **common)  # fmt: skip
        mc.accessible_volume = float(abs(np.linalg.det(cells[0])))
    elif cell_move is not None:
        cls = Isotension if c["isotension"] else Isobaric
        kw = {"external_stress": np.array(c["stress"])} if c["isotension"] else {}
        mc = cls(atoms, pressure=c["pressure"], default_displacement_move=disp, default_cell_move=cell_move, **kw, **common)
    else:
        mc = Canonical(atoms, default_displacement_move=None if c["composite_k"] > 1 else disp, **common)
        if c["composite_k"] > 1:
            mc.add_move(disp * c["composite_k"], criteria=CanonicalCriteria(), name="default_displacement_move")
    order = ["default_displacement_move"] + (["default_cell_move"] if "cell" in c["kinds"] else []) + (["default_exchange_move"] if "exchange" in c["kinds"] else [])
    assert list(mc.moves) == order
    kinds_in_order = [{"default_displacement_move": "disp", "default_cell_move": "cell", "default_exchange_move": "exchange"}[k] for k in order]
    for name, kind in zip(order, kinds_in_order):
        mc.moves[name].probability = c["probabilities"][c["kinds"].index(kind)]
    mc.run(1)
    tr1 = {k: v.copy() for k, v in mc.last_trace.items()}
    mc.run(1)  # a second launch: everything that must persist between launches does
    tr2 = mc.last_trace
    trace = {k: np.concatenate([tr1[k], tr2[k]], axis=1) for k in ("move", "accepted", "energy")}
    return mc, trace, kinds_in_order


def _run_oracle(c, pos, cells, counts, pbc, n, n_max, kinds_in_order, r):
    lj = c["pot"] == "lj"
    par = LJParams(sigma=2.3, epsilon=0.2, rc=6.0) if lj else emt_params("Cu")
    gc = "exchange" in c["kinds"]
    rep = capi.Replica(
        positions=pos[r].copy(), cell=cells[r], pbc=tuple(int(b) for b in pbc), n_atoms=int(counts[r]), temperature=c["temperature"],
        pressure=c["pressure"] if "cell" in c["kinds"] else 0.0, external_stress=np.array(c["stress"]) if c["isotension"] else None, mass=63.546, chemical_potential=c["mu"], exchange_mass=63.546,
        n_exchange=n if gc else 0, accessible_volume=float(abs(np.linalg.det(cells[0]))),
    )  # fmt: skip
    ld = np.full(n_max, -1, np.int32)
    ld[: counts[r]] = np.arange(counts[r])
    lx = ld.copy()
    specs = []
    for kind in kinds_in_order:
        p = c["probabilities"][c["kinds"].index(kind)]
        if kind == "disp":
            extra = ((OPS[c["extra_op"][0]], c["extra_op"][1]),) if c["extra_op"] else ()
            specs.append(capi.MoveSpec(capi.MOVE_DISPLACEMENT, OPS[c["disp_op"]], step=c["disp_step"], probability=p, labels=ld,
                                       repeat=c["composite_k"], extra_ops=extra))  # fmt: skip
        elif kind == "cell":
            axes = int(sum(int(b) << d for d, b in enumerate(c["iso_mask"]))) if c["cell_op"] == "Isotropic" else 0
            specs.append(capi.MoveSpec(capi.MOVE_CELL, CELL_OPS[c["cell_op"]], step=c["cell_max"], probability=p, axes=axes))
        else:
            specs.append(capi.MoveSpec(capi.MOVE_EXCHANGE, capi.OP_TRANSLATION, probability=p, labels=lx))
    tr = capi.mc_run(par, rep, specs, 4242 + c["seed"], r, 0, 2 * c["n_attempts"])
    return rep, tr, ld, lx


@pytest.mark.parametrize("seed", range(OFFSET, OFFSET + 28 * SCALE))
def test_random_case_matches_oracle(cuda_lib, seed):
    c = _case(seed)
    pos, cells, counts, pbc, n, n_max = _system(c)
    mc, trace, kinds_in_order = _run_gpu(c, pos, cells, counts, pbc, n, n_max)
    out = mc.download()
    for r in range(c["R"]):
        rep, ref, ld, lx = _run_oracle(c, pos, cells, counts, pbc, n, n_max, kinds_in_order, r)
        assert np.array_equal(trace["move"][r], ref["move"]), c
        assert np.array_equal(trace["accepted"][r], ref["accepted"]), f"accept/reject sequence differs, replica {r}, case {c}"
        np.testing.assert_allclose(trace["energy"][r], ref["energy"], rtol=1e-10, atol=1e-10, err_msg=str(c))
        m = rep.n_atoms
        assert out.n_atoms[r] == m
        np.testing.assert_allclose(out.positions[r, :m], rep.positions[:m], rtol=0, atol=1e-10, err_msg=str(c))
        np.testing.assert_allclose(out.cell[r], rep.cell, rtol=0, atol=1e-12, err_msg=str(c))
        if "exchange" in c["kinds"]:
            assert mc.number_of_exchange_particles[r] == rep.n_exchange
            assert np.array_equal(mc.labels("default_exchange_move")[r, :m], lx[:m])
            assert np.array_equal(mc.labels("default_displacement_move")[r, :m], ld[:m])
    assert int(mc.batch.status.max().item()) == 0, c


# ---------------------------------------------------------------------------------------------------------------------------
# second family: grand-canonical tables with molecules, composite exchange moves and the label groups they leave behind
# ---------------------------------------------------------------------------------------------------------------------------
def _gc_case(seed):
    rng = np.random.default_rng(30_000 + seed)
    c = {"seed": seed, "R": 3, "n0": int(rng.integers(10, 26)), "L": float(rng.uniform(11.0, 14.0)), "n_max": 96}
    c["pbc"] = bool(rng.random() < 0.6)
    c["eps"], c["sigma"] = float(rng.uniform(0.03, 0.08)), float(rng.uniform(2.3, 2.7))
    c["temperature"] = float(rng.uniform(300.0, 600.0))
    c["mu"] = float(rng.uniform(-0.6, -0.4))
    c["kind"] = str(rng.choice(["single", "molecule", "molecule", "composite", "composite"]))
    c["k"] = int(rng.integers(2, 4))  # atoms per molecule / sub-moves per composite
    c["rotate"] = bool(c["kind"] == "molecule" and rng.random() < 0.7)
    mol = np.zeros((c["k"], 3))
    for q in range(1, c["k"]):
        v = rng.normal(size=3)
        mol[q] = mol[q - 1] + v / np.linalg.norm(v) * rng.uniform(2.6, 3.0)
    c["mol"] = mol.tolist()
    c["deletable"] = bool(rng.random() < 0.5)  # the initial atoms carry exchange labels of their own
    c["bias"] = float(rng.choice([0.5, 0.5, 0.35, 0.65])) if c["kind"] != "composite" else 0.5
    c["p_disp"] = float(rng.uniform(0.3, 0.6))
    c["disp_step"] = float(rng.uniform(0.05, 0.25))
    c["n_attempts"] = 160
    # label groups of the caller's own (pairs of neighbouring atoms share a label in BOTH moves): a displacement translates the
    # pair, a deletion takes both atoms out -- also next to single-atom insertions (moves/exchange.py:127, moves/displacement.py:164)
    c["pairs"] = bool(seed % 24 >= 16)
    return c


@pytest.mark.parametrize("seed", range(OFFSET, OFFSET + 24 * SCALE))
def test_random_grand_canonical_case_matches_oracle(cuda_lib, seed):
    from quansino_b200 import BatchedAtoms, LennardJones, operations
    from quansino_b200.mc import GrandCanonical
    from quansino_b200.mc.criteria import GrandCanonicalCriteria
    from quansino_b200.moves import DisplacementMove, ExchangeMove

    c = _gc_case(seed)
    rng = np.random.default_rng(40_000 + seed)
    R, n0, n_max, L = c["R"], c["n0"], c["n_max"], c["L"]
    cell = np.eye(3) * L
    pbc = (c["pbc"],) * 3
    pos = np.zeros((R, n_max, 3))
    grid = np.array([[i, j, k] for i in range(3) for j in range(3) for k in range(3)], dtype=float)
    for r in range(R):
        pos[r, :n0] = ((grid[rng.permutation(27)[:n0]] + 0.5) / 3.0 + rng.normal(scale=0.02, size=(n0, 3))) * L
    calc = LennardJones(sigma=c["sigma"], epsilon=c["eps"], rc=6.0)
    par = LJParams(sigma=c["sigma"], epsilon=c["eps"], rc=6.0)
    mass1 = 195.084
    molecular = c["kind"] == "molecule"
    mol = np.array(c["mol"]) if molecular else np.zeros((1, 3))
    lab_x = np.arange(n0) if c["deletable"] else np.full(n0, -1)
    lab_d = np.arange(n0)
    if c["pairs"]:
        lab_d = lab_x = np.arange(n0) // 2
        c["deletable"] = True
    n_ex0 = int(lab_x.max() + 1) if c["deletable"] else 0
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n0, dtype=np.int32), "Pt", pbc, calc)
    op = operations.TranslationRotation() if c["rotate"] else operations.Translation()
    mc = GrandCanonical(
        atoms, exchange_atoms=("Pt", mol), temperature=c["temperature"], chemical_potential=c["mu"],
        number_of_exchange_particles=n_ex0, max_cycles=c["n_attempts"], seed=777 + seed, trace=True, resync_interval=0,
        default_displacement_move=DisplacementMove(lab_d.copy(), operations.Ball(c["disp_step"])),
    )  # fmt: skip
    if c["kind"] == "composite":
        exch = ExchangeMove(lab_x.copy())
        for _ in range(c["k"] - 1):
            exch = exch + ExchangeMove(lab_x.copy())
    else:
        exch = ExchangeMove(lab_x.copy(), op, bias_towards_insert=c["bias"])
    mc.add_move(exch, criteria=GrandCanonicalCriteria(), name="default_exchange_move")
    mc.accessible_volume = L**3
    mc.moves["default_displacement_move"].probability = c["p_disp"]
    mc.moves["default_exchange_move"].probability = 1 - c["p_disp"]
    mc.run(1)
    tr1 = {k: v.copy() for k, v in mc.last_trace.items()}
    mc.run(1)
    tr = {k: np.concatenate([tr1[k], mc.last_trace[k]], axis=1) for k in ("move", "accepted", "energy", "moved")}
    out, n_ex = mc.download(), mc.number_of_exchange_particles
    labels_d, labels_x = mc.labels("default_displacement_move"), mc.labels("default_exchange_move")
    changes: dict[int, int] = {}
    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cell, pbc=tuple(int(b) for b in pbc), n_atoms=n0, temperature=c["temperature"],
                           mass=mass1, chemical_potential=c["mu"], exchange_mass=mass1 * len(mol), n_exchange=n_ex0,
                           accessible_volume=L**3, exchange_mol=mol if molecular else None)  # fmt: skip
        ld = np.full(n_max, -1, np.int32)
        ld[:n0] = lab_d
        lx = np.full(n_max, -1, np.int32)
        lx[:n0] = lab_x
        specs = [
            capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=c["disp_step"], probability=c["p_disp"], labels=ld),
            capi.MoveSpec(capi.MOVE_EXCHANGE, capi.OP_TRANSLATION_ROTATION if c["rotate"] else capi.OP_TRANSLATION, probability=1 - c["p_disp"],
                          labels=lx, repeat=c["k"] if c["kind"] == "composite" else 1, bias_insert=c["bias"], chain=8 if c["pairs"] else 0),
        ]  # fmt: skip
        ref = capi.mc_run(par, rep, specs, 777 + seed, r, 0, 2 * c["n_attempts"])
        assert np.array_equal(tr["move"][r], ref["move"]), c
        assert np.array_equal(tr["accepted"][r], ref["accepted"]), f"accept/reject sequence differs, replica {r}, case {c}"
        assert np.array_equal(tr["moved"][r], ref["moved"]), c
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-10, err_msg=str(c))
        m = rep.n_atoms
        assert out.n_atoms[r] == m and n_ex[r] == rep.n_exchange, c
        assert np.array_equal(labels_d[r, :m], ld[:m]) and np.array_equal(labels_x[r, :m], lx[:m]), c
        np.testing.assert_allclose(out.positions[r, :m], rep.positions[:m], rtol=0, atol=1e-10, err_msg=str(c))
        d_n = np.diff(np.concatenate([[n0], ref["n_atoms"]]))
        for v, cnt in zip(*np.unique(d_n, return_counts=True)):
            changes[int(v)] = changes.get(int(v), 0) + int(cnt)
    assert int(mc.batch.status.max().item()) == 0, c
    if seed < 18:  # liveness of the committed seeds: something was inserted or deleted (a scaled hunt may draw a case where nothing is accepted; seed 85 is one)
        assert sum(v for k, v in changes.items() if k != 0) > 0, (c, changes)


# ---------------------------------------------------------------------------------------------------------------------------
# third family: label groups in arbitrary order with the rigid-body operations, hybrid NPT composites, Hamiltonian moves
# ---------------------------------------------------------------------------------------------------------------------------
def _misc_case(seed):
    rng = np.random.default_rng(50_000 + seed)
    c = {"seed": seed, "R": 3, "kind": str(rng.choice(["groups", "groups", "hybrid", "hmc", "hmc"]))}
    c["pot"] = str(rng.choice(["emt", "emt", "lj"]))
    c["repeat"] = int(rng.choice([3, 3, 4]))
    c["slab"] = bool(c["kind"] != "hybrid" and rng.random() < 0.3)
    c["sheared"] = bool(c["kind"] == "hmc" and not c["slab"] and rng.random() < 0.5)
    c["temperature"] = float(rng.choice([300.0, 800.0, 1500.0]))
    c["step"] = float(rng.uniform(0.04, 0.15))
    c["group_op"] = str(rng.choice(["Rotation", "Translation", "TranslationRotation"]))
    c["k"] = int(rng.integers(2, 5))
    c["cell_max"] = float(rng.uniform(0.003, 0.01))
    c["p"] = float(rng.uniform(0.3, 0.7))
    c["dt_fs"] = float(rng.uniform(0.5, 3.0))
    c["n_steps"] = int(rng.integers(3, 16))
    c["n_attempts"] = 3 if c["kind"] == "hmc" else 70
    return c


@pytest.mark.parametrize("seed", range(OFFSET, OFFSET + 18 * SCALE))
def test_random_groups_hybrid_hmc_case_matches_oracle(cuda_lib, seed):
    from quansino_b200 import EMT, BatchedAtoms, LennardJones, operations
    from quansino_b200.integrators import Verlet
    from quansino_b200.integrators.displacement import fs
    from quansino_b200.mc import Canonical, HamiltonianCanonical, Isobaric
    from quansino_b200.mc.criteria import IsobaricCriteria
    from quansino_b200.moves import CellMove, DisplacementMove, HamiltonianDisplacementMove

    c = _misc_case(seed)
    rng = np.random.default_rng(60_000 + seed)
    pos0, cell0 = fcc_cubic("Cu", c["repeat"])
    n, R = len(pos0), c["R"]
    cells = np.repeat(cell0[None], R, axis=0)
    pbc = (True, True, not c["slab"])
    pos = np.zeros((R, n, 3))
    for r in range(R):
        p = rattle(pos0, 0.04, 7000 + 10 * seed + r)
        if c["sheared"]:
            S = np.eye(3) + rng.uniform(-0.07, 0.07, (3, 3)) * (1 - np.eye(3))
            cells[r] = S @ cell0
            p = np.linalg.solve(cell0.T, p.T).T @ cells[r]
        if c["slab"]:
            cells[r, 2, 2] = cell0[2, 2] * 1.7
        pos[r] = p
    lj = c["pot"] == "lj"
    calc = LennardJones(sigma=2.3, epsilon=0.2, rc=6.0) if lj else EMT()
    par = LJParams(sigma=2.3, epsilon=0.2, rc=6.0) if lj else emt_params("Cu")
    atoms = BatchedAtoms(pos.copy(), cells, np.full(R, n, dtype=np.int32), "Cu", pbc, calc)
    common = dict(temperature=c["temperature"], seed=999 + seed, trace=True, resync_interval=0)
    if c["kind"] == "groups":
        # groups of 1 .. 4 atoms scattered over the index range, label VALUES in arbitrary order, some atoms immovable (-1)
        ids = rng.permutation(np.repeat(np.arange(n // 2), 2))[:n]
        values = rng.permutation(3 * n)[: n // 2]
        lab = np.where(rng.random(n) < 0.1, -1, values[ids]).astype(np.int32)
        mc = Canonical(atoms, max_cycles=c["n_attempts"], default_displacement_move=DisplacementMove(lab, operations.Ball(c["step"])), **common)
        mc.add_move(DisplacementMove(lab, getattr(operations, c["group_op"])()), name="rigid", probability=1.0)
        mc.moves["default_displacement_move"].probability = c["p"]
        mc.moves["rigid"].probability = 1 - c["p"]
        op2 = {"Rotation": capi.OP_ROTATION, "Translation": capi.OP_TRANSLATION, "TranslationRotation": capi.OP_TRANSLATION_ROTATION}[c["group_op"]]
        specs = lambda: [capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=c["step"], probability=c["p"], labels=lab.copy()),
                         capi.MoveSpec(capi.MOVE_DISPLACEMENT, op2, probability=1 - c["p"], labels=lab.copy())]  # noqa: E731  # fmt: skip
        state = {}
    elif c["kind"] == "hybrid":
        lab = np.arange(n)
        mc = Isobaric(atoms, pressure=300.0 * BAR, max_cycles=c["n_attempts"], default_displacement_move=DisplacementMove(lab, operations.Ball(c["step"])), **common)
        hybrid = DisplacementMove(lab, operations.Ball(0.5 * c["step"])) * c["k"] + CellMove(operations.IsotropicDeformation(c["cell_max"]))
        mc.add_move(hybrid, criteria=IsobaricCriteria(), name="hybrid", probability=1.0)
        mc.moves["default_displacement_move"].probability = c["p"]
        mc.moves["hybrid"].probability = 1 - c["p"]
        labs = np.arange(n, dtype=np.int32)
        specs = lambda: [capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=c["step"], probability=c["p"], labels=labs.copy()),
                         capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.5 * c["step"], probability=1 - c["p"], labels=labs.copy(), repeat=c["k"], chain=3),
                         capi.MoveSpec(capi.MOVE_CELL, capi.OP_ISOTROPIC, step=c["cell_max"], probability=1 - c["p"], chain=2)]  # noqa: E731  # fmt: skip
        state = {"pressure": 300.0 * BAR}
    else:
        mc = HamiltonianCanonical(atoms, max_cycles=1, steps_per_launch=1,
                                  default_displacement_move=HamiltonianDisplacementMove(operation=Verlet(dt=c["dt_fs"], max_steps=c["n_steps"])), **common)  # fmt: skip
        specs = lambda: [capi.MoveSpec(capi.MOVE_HMC, capi.OP_VERLET, dt=c["dt_fs"] * fs, n_steps=c["n_steps"])]  # noqa: E731
        state = {"mass": 63.546}
    traces = []
    launches = c["n_attempts"] if c["kind"] == "hmc" else 2
    for _ in range(launches):
        mc.run(1)
        traces.append({k: v.copy() for k, v in mc.last_trace.items()})
    tr = {k: np.concatenate([t[k] for t in traces], axis=1) for k in ("accepted", "energy")}
    out = mc.download()
    total = launches * (1 if c["kind"] == "hmc" else c["n_attempts"])
    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cells[r], pbc=tuple(int(b) for b in pbc), temperature=c["temperature"],
                           momenta=np.zeros((n, 3)) if c["kind"] == "hmc" else None, **state)  # fmt: skip
        ref = capi.mc_run(par, rep, specs(), 999 + seed, r, 0, total)
        assert np.array_equal(tr["accepted"][r], ref["accepted"]), f"accept/reject sequence differs, replica {r}, case {c}"
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-10, err_msg=str(c))
        np.testing.assert_allclose(out.positions[r], rep.positions, rtol=0, atol=1e-9 if c["kind"] == "hmc" else 1e-10, err_msg=str(c))
        np.testing.assert_allclose(out.cell[r], rep.cell, rtol=1e-13, atol=0, err_msg=str(c))
        if c["kind"] == "hmc":
            np.testing.assert_allclose(out.momenta[r], rep.momenta, rtol=0, atol=1e-9, err_msg=str(c))
    assert int(mc.batch.status.max().item()) == 0, c
