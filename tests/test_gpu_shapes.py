"""Every atom-capacity class of the sweep kernels (32 ... 1024 atoms per replica) against the CPU oracle: the 32-atom cell needs
up to seven periodic images per neighbour, the 500-atom cell is BASELINE.json's isobaric configuration, 864 atoms fill the
largest class.  Also a batch of several waves of CTAs."""

import os

import numpy as np
import pytest

from oracle import capi
from tests.helpers import cu_replicas, oracle_par, oracle_run

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_variant")]

BAR = 6.241509125883258e-07


@pytest.mark.parametrize("repeat,n_attempts", [(2, 200), (4, 120), (6, 60)])
def test_canonical_capacity_classes(cuda_lib, repeat, n_attempts):
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove
    import os

    if repeat == 2 and os.environ.get("QMC_FORCE_KERNEL") == "block":
        pytest.skip("the CTA-per-replica kernel is sized for cells of >= 64 atoms (never selected below 256 / on request only)")
    R = 3
    pos, cell, n = cu_replicas(repeat, R, seed=40 + repeat)
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    mc = Canonical(atoms, temperature=400.0, seed=700 + repeat, trace=True, max_cycles=n_attempts, resync_interval=0,
                   neighbour_rows=os.environ.get("QMC_FORCE_KERNEL") == "rows",
                   default_displacement_move=DisplacementMove(np.arange(n), operations.Ball(0.12)))  # fmt: skip
    e0 = mc.batch.energy_full().cpu().numpy().copy()
    mc.run(1)
    tr, out = mc.last_trace, mc.download()
    par = oracle_par(mc.batch.calc)
    reps, otr = oracle_run(par, pos, np.full(R, n), cell, (1, 1, 1), [dict(type=capi.MOVE_DISPLACEMENT, op=capi.OP_BALL, step=0.12)],
                           700 + repeat, n_attempts, range(R), temperature=400.0)  # fmt: skip
    for r in range(R):
        ref0 = capi.energy(par, pos[r], cell, (1, 1, 1))
        assert abs(e0[r] - ref0["energy"]) <= 1e-10 * abs(ref0["energy"])
        assert np.array_equal(tr["accepted"][r], otr[r]["accepted"]), f"accept/reject sequence differs for replica {r}"
        assert np.array_equal(tr["moved"][r], otr[r]["moved"])
        np.testing.assert_allclose(tr["energy"][r], otr[r]["energy"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out.positions[r], reps[r].positions, rtol=0, atol=1e-12)
    assert int(mc.batch.status.max().item()) == 0


def test_isobaric_500_atoms(cuda_lib):
    """BASELINE.json configuration C3 at test size: 500-atom cells, displacement and cell moves (cell probability raised)."""
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import Isobaric
    from quansino_b200.moves import CellMove, DisplacementMove

    R, n_attempts, p_cell = 3, 60, 0.12
    pos, cell, n = cu_replicas(5, R, seed=55)
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    mc = Isobaric(atoms, temperature=300.0, pressure=1.0 * BAR, seed=808, trace=True, max_cycles=n_attempts, resync_interval=0,
                  default_displacement_move=DisplacementMove(np.arange(n), operations.Ball(0.1)),
                  default_cell_move=CellMove(operations.IsotropicDeformation(0.01)))  # fmt: skip
    mc.moves["default_cell_move"].probability = p_cell
    mc.moves["default_displacement_move"].probability = 1.0 - p_cell
    mc.run(1)
    tr, out = mc.last_trace, mc.download()
    par = oracle_par(mc.batch.calc)
    moves = [dict(type=capi.MOVE_DISPLACEMENT, op=capi.OP_BALL, step=0.1, probability=1.0 - p_cell),
             dict(type=capi.MOVE_CELL, op=capi.OP_ISOTROPIC, step=0.01, probability=p_cell)]  # fmt: skip
    reps, otr = oracle_run(par, pos, np.full(R, n), cell, (1, 1, 1), moves, 808, n_attempts, range(R), temperature=300.0, pressure=1.0 * BAR)
    n_cell = 0
    for r in range(R):
        assert np.array_equal(tr["move"][r], otr[r]["move"])
        assert np.array_equal(tr["accepted"][r], otr[r]["accepted"]), f"accept/reject sequence differs for replica {r}"
        np.testing.assert_allclose(tr["energy"][r], otr[r]["energy"], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(out.positions[r], reps[r].positions, rtol=0, atol=1e-10)
        np.testing.assert_allclose(out.cell[r], reps[r].cell, rtol=1e-13)
        n_cell += int((otr[r]["move"] == 1).sum())
    assert n_cell >= 6


def test_isobaric_c3_full_step_32_replicas(cuda_lib):
    """BASELINE.json configuration C3 as the bench runs it, at a size the oracle finishes in seconds: 32 replicas of the 500-atom
    cell, ONE full step of the default move table (max_cycles = 500 attempts, cell move with the default probability 1 / (N + 1),
    IsotropicDeformation 0.05), every attempt of every replica against the C oracle (`mc/isobaric.py:117-147`,
    `mc/criteria.py:146-172`)."""
    import os

    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import Isobaric
    from quansino_b200.moves import CellMove, DisplacementMove

    R = 32
    pos, cell, n = cu_replicas(5, R, seed=56)
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    mc = Isobaric(atoms, temperature=300.0, pressure=1.0 * BAR, seed=909, trace=True, resync_interval=0,
                  default_displacement_move=DisplacementMove(np.arange(n), operations.Ball(0.1)),
                  default_cell_move=CellMove(operations.IsotropicDeformation(0.05)))  # fmt: skip
    n_attempts = mc.max_cycles
    assert n_attempts == n  # `mc/canonical.py:70-72`: max_cycles defaults to the number of atoms; the cell move shares them
    p_disp, p_cell = mc.moves["default_displacement_move"].probability, mc.moves["default_cell_move"].probability
    mc.run(1)
    tr, out = mc.last_trace, mc.download()
    par = oracle_par(mc.batch.calc)
    moves = [dict(type=capi.MOVE_DISPLACEMENT, op=capi.OP_BALL, step=0.1, probability=p_disp),
             dict(type=capi.MOVE_CELL, op=capi.OP_ISOTROPIC, step=0.05, probability=p_cell)]  # fmt: skip
    reps, otr = oracle_run(par, pos, np.full(R, n), cell, (1, 1, 1), moves, 909, n_attempts, range(R), threads=min(16, os.cpu_count() or 1),
                           temperature=300.0, pressure=1.0 * BAR)  # fmt: skip
    n_cell = n_cell_acc = 0
    for r in range(R):
        assert np.array_equal(tr["move"][r], otr[r]["move"])
        assert np.array_equal(tr["accepted"][r], otr[r]["accepted"]), f"accept/reject sequence differs for replica {r}"
        np.testing.assert_allclose(tr["energy"][r], otr[r]["energy"], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(out.positions[r], reps[r].positions, rtol=0, atol=1e-10)
        np.testing.assert_allclose(out.cell[r], reps[r].cell, rtol=1e-13)
        cells = otr[r]["move"] == 1
        n_cell += int(cells.sum())
        n_cell_acc += int((otr[r]["accepted"][cells] == 1).sum())
    assert n_cell >= 15, n_cell
    assert int(mc.batch.status.max().item()) == 0


def test_several_waves_of_replicas(cuda_lib):
    """10 000 replicas = 2.4 waves of CTAs: every replica is written exactly once, sampled replicas match the oracle."""
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove
    import os

    R = 10000
    pos, cell, n = cu_replicas(3, R, seed=3)
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    mc = Canonical(atoms, temperature=300.0, seed=4242, trace=True, max_cycles=n, resync_interval=0,
                   default_displacement_move=DisplacementMove(np.arange(n), operations.Ball(0.1)))  # fmt: skip
    mc.run(1)
    tr, out = mc.last_trace, mc.download()
    cnt = mc.counters()
    assert np.array_equal(cnt[:, 0, 0], np.full(R, n)) and np.array_equal(cnt[:, 0, 1], tr["accepted"].sum(axis=1))
    ids = [0, 4143, 4144, 8287, 8288, 9999]
    par = oracle_par(mc.batch.calc)
    reps, otr = oracle_run(par, pos[ids], np.full(len(ids), n), cell, (1, 1, 1), [dict(type=capi.MOVE_DISPLACEMENT, op=capi.OP_BALL, step=0.1)],
                           4242, n, ids, temperature=300.0)  # fmt: skip
    for k, r in enumerate(ids):
        assert np.array_equal(tr["accepted"][r], otr[k]["accepted"])
        np.testing.assert_allclose(out.positions[r], reps[k].positions, rtol=0, atol=1e-12)
    tracked = mc.batch.energy.cpu().numpy().copy()
    np.testing.assert_allclose(tracked, mc.batch.energy_full().cpu().numpy(), rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("case", ["orthorhombic", "slab", "unwrapped"])
def test_cell_geometries(cuda_lib, case):
    """Unequal cell edges, a slab (pbc T, T, F) and atoms that drifted several cells away from the box: positions are never
    wrapped by the reference (moves/displacement.py:109-141), the images must still be found."""
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove
    from quansino_b200.systems import fcc_cubic, rattle
    import os

    R, n_attempts = 3, 150
    pos0, cell0 = fcc_cubic("Cu", 3)
    L = cell0[0, 0]
    pbc = (True, True, True)
    if case == "orthorhombic":  # two cubes stacked along z: 216 atoms in a 10.83 x 10.83 x 21.66 cell
        pos0 = np.concatenate([pos0, pos0 + np.array([0.0, 0.0, L])])
        cell = np.diag([L, L, 2 * L])
    elif case == "slab":
        cell = np.diag([L, L, L + 12.0])
        pbc = (True, True, False)
    else:
        cell = cell0
    n = len(pos0)
    pos = np.stack([rattle(pos0, 0.05, 60 + r) for r in range(R)])
    if case == "unwrapped":
        rng = np.random.default_rng(8)
        pos += rng.integers(-3, 4, size=(R, n, 3)) * L  # every atom somewhere within +-3 cells
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), "Cu", pbc, EMT())
    mc = Canonical(atoms, temperature=500.0, seed=99, trace=True, max_cycles=n_attempts, resync_interval=0,
                   neighbour_rows=os.environ.get("QMC_FORCE_KERNEL") == "rows",
                   default_displacement_move=DisplacementMove(np.arange(n), operations.Sphere(0.1)))  # fmt: skip
    mc.run(1)
    tr, out = mc.last_trace, mc.download()
    par = oracle_par(mc.batch.calc)
    pbc_i = tuple(int(b) for b in pbc)
    reps, otr = oracle_run(par, pos, np.full(R, n), cell, pbc_i, [dict(type=capi.MOVE_DISPLACEMENT, op=capi.OP_SPHERE, step=0.1)], 99,
                           n_attempts, range(R), temperature=500.0)  # fmt: skip
    for r in range(R):
        assert np.array_equal(tr["accepted"][r], otr[r]["accepted"]), f"accept/reject sequence differs for replica {r}"
        assert np.array_equal(tr["moved"][r], otr[r]["moved"])
        np.testing.assert_allclose(tr["energy"][r], otr[r]["energy"], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(out.positions[r], reps[r].positions, rtol=0, atol=1e-10)
    if case == "unwrapped":  # the same configuration wrapped back has the same energy
        wrapped = BatchedAtoms(pos - np.floor(pos / L) * L, cell, np.full(R, n, dtype=np.int32), "Cu", pbc, EMT())
        from quansino_b200 import ReplicaBatch

        e_w = ReplicaBatch(wrapped, temperature=500.0).energy_full().cpu().numpy()
        e_u = ReplicaBatch(BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), "Cu", pbc, EMT()), temperature=500.0).energy_full().cpu().numpy()
        np.testing.assert_allclose(e_u, e_w, rtol=1e-11)


@pytest.mark.parametrize("build_lanes", [None, 32])
@pytest.mark.parametrize("case", ["cell32", "cell864", "cell2048", "slab", "unwrapped", "lj_cluster"])
def test_neighbour_list_forces_on_other_shapes(cuda_lib, monkeypatch, case, build_lanes):
    """The Hamiltonian kernels' forces and energies (Verlet-list build, image shifts, lanes per atom) against the oracle;
    ``build_lanes`` 32 forces the split list build with its bounding-box test onto every shape (None: the automatic choice)."""
    if build_lanes is not None:
        monkeypatch.setenv("QMC_HMC_BUILD_LANES", str(build_lanes))
    from quansino_b200 import EMT, BatchedAtoms, LennardJones, ReplicaBatch
    from quansino_b200.systems import fcc_cubic, octahedron, rattle
    from oracle.potentials import LJParams, emt_params

    if os.environ.get("QMC_FORCE_KERNEL") not in (None, "warp"):
        pytest.skip("the force kernels do not depend on the sweep-kernel variant")
    pbc, calc, par, symbol = (True, True, True), EMT(), emt_params("Cu"), "Cu"
    if case == "lj_cluster":
        pos0, cell = octahedron("Pt", 4, 10.0)
        pbc, calc, par, symbol = (False, False, False), LennardJones(sigma=2.5, epsilon=0.1), LJParams(sigma=2.5, epsilon=0.1), "Pt"
    else:
        repeat = {"cell32": 2, "cell864": 6, "cell2048": 8}.get(case, 3)
        pos0, cell = fcc_cubic("Cu", repeat)
    R = 1 if case == "cell2048" else 2
    L = cell[0, 0]
    if case == "slab":
        cell = np.diag([L, L, L + 12.0])
        pbc = (True, True, False)
    pos = np.stack([rattle(pos0, 0.05, 70 + r) for r in range(R)])
    if case == "unwrapped":
        pos += np.random.default_rng(9).integers(-3, 4, size=pos.shape) * L
    n = pos.shape[1]
    batch = ReplicaBatch(BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), symbol, pbc, calc), temperature=300.0)
    f = batch.forces_full().cpu().numpy().transpose(0, 2, 1)
    e = batch.energy.cpu().numpy()
    for r in range(R):
        ref = capi.energy(par, pos[r], cell, tuple(int(b) for b in pbc), want_forces=True)
        np.testing.assert_allclose(f[r], ref["forces"], rtol=1e-9, atol=1e-10)
        assert abs(e[r] - ref["energy"]) <= 1e-10 * abs(ref["energy"]) + 1e-12
    assert int(batch.status.max().item()) == 0


def test_lennard_jones_isobaric_and_hamiltonian(cuda_lib):
    """The pair potential through the cell-move (full recompute) and Hamiltonian (forces + Verlet) paths: periodic LJ solid
    with two images per axis, and the non-periodic Pt44 cluster."""
    from quansino_b200 import BatchedAtoms, LennardJones, operations
    from quansino_b200.integrators import Verlet
    from quansino_b200.integrators.displacement import fs
    from quansino_b200.mc import HamiltonianCanonical, Isobaric
    from quansino_b200.moves import CellMove, DisplacementMove, HamiltonianDisplacementMove
    from quansino_b200.systems import fcc_cubic, octahedron, rattle
    from oracle.potentials import LJParams

    par = LJParams(sigma=2.5, epsilon=0.1)
    calc = LennardJones(sigma=2.5, epsilon=0.1)
    # --- isobaric, periodic: fcc with a = 3.92 (Pt spacing), 108 atoms, L = 11.76 < 2 x 7.5
    R, n_att, p_cell = 3, 120, 0.15
    pos0, cell = fcc_cubic("Pt", 3)
    n = len(pos0)
    pos = np.stack([rattle(pos0, 0.05, 80 + r) for r in range(R)])
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), "Pt", (True, True, True), calc)
    mc = Isobaric(atoms, temperature=300.0, pressure=500.0 * BAR, seed=31, trace=True, max_cycles=n_att, resync_interval=0,
                  default_displacement_move=DisplacementMove(np.arange(n), operations.Ball(0.1)),
                  default_cell_move=CellMove(operations.IsotropicDeformation(0.01)))  # fmt: skip
    mc.moves["default_cell_move"].probability = p_cell
    mc.moves["default_displacement_move"].probability = 1.0 - p_cell
    mc.run(1)
    tr, out = mc.last_trace, mc.download()
    moves = [dict(type=capi.MOVE_DISPLACEMENT, op=capi.OP_BALL, step=0.1, probability=1.0 - p_cell),
             dict(type=capi.MOVE_CELL, op=capi.OP_ISOTROPIC, step=0.01, probability=p_cell)]  # fmt: skip
    reps, otr = oracle_run(par, pos, np.full(R, n), cell, (1, 1, 1), moves, 31, n_att, range(R), temperature=300.0, pressure=500.0 * BAR,
                           mass=195.084)  # fmt: skip
    for r in range(R):
        assert np.array_equal(tr["move"][r], otr[r]["move"]) and (otr[r]["move"] == 1).sum() > 5
        assert np.array_equal(tr["accepted"][r], otr[r]["accepted"]), f"accept/reject sequence differs for replica {r}"
        np.testing.assert_allclose(tr["energy"][r], otr[r]["energy"], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(out.cell[r], reps[r].cell, rtol=1e-13)
    if os.environ.get("QMC_FORCE_KERNEL") not in (None, "warp"):
        return  # the Hamiltonian kernels do not depend on the sweep-kernel variant
    # --- Hamiltonian, non-periodic cluster
    pos0, cell = octahedron("Pt", 4, 10.0)
    n = len(pos0)
    R, n_attempts, n_steps = 2, 3, 10
    pos = np.stack([rattle(pos0, 0.02, 90 + r) for r in range(R)])
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), "Pt", (False, False, False), calc)
    hm = HamiltonianCanonical(atoms, temperature=np.array([300.0, 700.0]), max_cycles=1, seed=17, trace=True, resync_interval=0,
                              steps_per_launch=n_attempts,
                              default_displacement_move=HamiltonianDisplacementMove(operation=Verlet(dt=2.0, max_steps=n_steps)))  # fmt: skip
    hm.run(n_attempts)
    tr, out = hm.last_trace, hm.download()
    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cell, pbc=(0, 0, 0), temperature=float([300.0, 700.0][r]), mass=195.084,
                           momenta=np.zeros((n, 3)))  # fmt: skip
        ref = capi.mc_run(par, rep, [capi.MoveSpec(capi.MOVE_HMC, capi.OP_VERLET, dt=2.0 * fs, n_steps=n_steps)], 17, r, 0, n_attempts)
        assert np.array_equal(tr["accepted"][r], ref["accepted"])
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(out.positions[r], rep.positions, rtol=0, atol=1e-10)
        np.testing.assert_allclose(out.momenta[r], rep.momenta, rtol=0, atol=1e-10)


def test_tiny_and_empty_replicas(cuda_lib):
    """Replicas with 0, 1, 2 and 3 atoms in one batch (an empty replica fails every move; a single atom has energy -E0 and
    accepts everything), canonical and grand-canonical, against the oracle."""
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove

    n_max, n_att = 32, 60
    cell = np.diag([14.0, 15.0, 16.0])
    n_atoms = np.array([0, 1, 2, 3], dtype=np.int32)
    R = len(n_atoms)
    rng = np.random.default_rng(12)
    pos = np.zeros((R, n_max, 3))
    for r in range(R):
        pos[r, : n_atoms[r]] = 6.0 + rng.normal(scale=1.2, size=(n_atoms[r], 3))  # a loose clump: pairs inside the cutoff
    atoms = BatchedAtoms(pos.copy(), cell, n_atoms, "Cu", (True, True, True), EMT())
    labels = np.tile(np.arange(n_max), (R, 1))
    mc = Canonical(atoms, temperature=2000.0, seed=64, trace=True, max_cycles=n_att, resync_interval=0,
                   default_displacement_move=DisplacementMove(labels, operations.Ball(0.3)))  # fmt: skip
    e0 = mc.batch.energy_full().cpu().numpy().copy()
    mc.run(1)
    tr, out, par = mc.last_trace, mc.download(), oracle_par(EMT())
    for r in range(R):
        n = int(n_atoms[r])
        ref0 = capi.energy(par, pos[r, :n], cell, (1, 1, 1)) if n else {"energy": 0.0}
        assert abs(e0[r] - ref0["energy"]) <= 1e-10 * abs(ref0["energy"]) + 1e-14
        rep = capi.Replica(positions=pos[r].copy(), cell=cell, n_atoms=n, temperature=2000.0)
        lab = labels[r].astype(np.int32).copy()
        lab[n:] = -1
        ref = capi.mc_run(par, rep, [capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.3, labels=lab)], 64, r, 0, n_att)
        assert np.array_equal(tr["accepted"][r], ref["accepted"]), f"replica with {n} atoms"
        assert np.array_equal(tr["moved"][r], ref["moved"])
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out.positions[r, :n], rep.positions[:n], rtol=0, atol=1e-12)
    assert np.all(tr["accepted"][0] == -1) and np.all(tr["accepted"][1] == 1)
    assert abs(e0[1] - 3.51) < 1e-12  # an isolated Cu atom sits at -E0 = +3.51 eV on ASE-EMT's energy scale


def test_limits_are_reported_where_the_move_is_registered(cuda_lib):
    """VERDICT r1 "weak" 10: a batch beyond the sweep kernels' 1024 atoms per replica refuses displacement / cell / exchange moves
    when they are added, not at the first launch; Hamiltonian moves take it."""
    from quansino_b200 import EMT, BatchedAtoms
    from quansino_b200.mc import Canonical, HamiltonianCanonical
    from quansino_b200.moves import DisplacementMove, HamiltonianDisplacementMove
    from quansino_b200.systems import fcc_cubic

    pos, cell = fcc_cubic("Cu", 8)
    atoms = BatchedAtoms.replicate(pos, cell, 1, calc=EMT())
    with pytest.raises(NotImplementedError, match="1024 atoms"):
        Canonical(atoms, temperature=300.0, default_displacement_move=DisplacementMove(np.arange(len(pos))))
    hm = HamiltonianCanonical(atoms, temperature=300.0, max_cycles=1, default_displacement_move=HamiltonianDisplacementMove())
    assert list(hm.moves) == ["default_displacement_move"]
