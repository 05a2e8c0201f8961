"""Randomised differential test: seeded random systems and move tables through the CUDA path and the C oracle.

Every case draws a potential (EMT / Lennard-Jones), a cell (cubic, slab, sheared, cluster in a box), ragged replica sizes, and a
move table out of the kinds the reference offers -- displacement (Ball / Box / Sphere, optionally a CompositeOperation or a
`move * k` composite), cell moves (the three deformation operations, Isobaric or Isotension criteria), plain exchange moves -- with
random step sizes and probabilities, runs it for a few hundred attempts over two launches, and compares every attempt of every
replica.  The point is breadth: combinations nobody wrote a dedicated test for."""

import numpy as np
import pytest

from oracle import capi
from oracle.potentials import LJParams, emt_params
from quansino_b200.systems import fcc_cubic, rattle

pytestmark = pytest.mark.gpu

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


@pytest.mark.parametrize("seed", range(28))
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
