"""General (triclinic) cells on the device (csrc/sweep_tri.cuh) against the C oracle: energies, displacement sweeps, the three
deformation operations (operations/cell.py:56-169) under IsobaricCriteria and IsotensionCriteria (mc/criteria.py:143-219).

Bar: identical accept / reject sequences, energies within 1e-10 relative, positions 1e-10 A, cells 1e-12 A."""

import numpy as np
import pytest

from oracle import capi
from oracle.potentials import LJParams, emt_params
from quansino_b200.systems import fcc_cubic, rattle

pytestmark = pytest.mark.gpu

BAR = 6.241509125883258e-07  # ase.units.bar


def _sheared(repeat, R, seed, strength=0.15, stdev=0.05, wander=0.0):
    """R rattled fcc Cu cells, each with its own shear (lower and upper off-diagonal terms): positions (R, n, 3), cells (R, 3, 3).
    ``wander``: atoms additionally displaced by whole cell vectors (positions are never wrapped by the reference)."""
    pos0, cell0 = fcc_cubic("Cu", repeat)
    n = len(pos0)
    rng = np.random.default_rng(seed)
    pos, cells = np.zeros((R, n, 3)), np.zeros((R, 3, 3))
    for r in range(R):
        S = np.eye(3) + rng.uniform(-strength, strength, (3, 3)) * (1 - np.eye(3))
        cells[r] = S @ cell0
        p = rattle(pos0, stdev, seed * 100 + r)
        frac = np.linalg.solve(cell0.T, p.T).T
        if wander:
            frac = frac + rng.integers(-2, 3, size=frac.shape) * (rng.random(frac.shape) < wander)
        pos[r] = frac @ cells[r]
    return pos, cells, n


@pytest.mark.parametrize("repeat,pot", [(2, "emt"), (3, "emt"), (4, "emt"), (5, "emt"), (3, "lj")])
def test_energies_in_general_cells_match_oracle(cuda_lib, repeat, pot):
    """Full energies (and the EMT caches) in sheared cells: 32 atoms (heights barely above the cutoff: up to eight images per
    pair) to 500 atoms; atoms that wandered by whole cell vectors."""
    from quansino_b200 import EMT, BatchedAtoms, LennardJones, ReplicaBatch

    R = 6
    pos, cells, n = _sheared(repeat, R, seed=40 + repeat, strength=0.12 if repeat > 2 else 0.04, wander=0.2)
    calc = EMT() if pot == "emt" else LennardJones(sigma=2.3, epsilon=0.2, rc=6.0)
    par = emt_params("Cu") if pot == "emt" else LJParams(sigma=2.3, epsilon=0.2, rc=6.0)
    if repeat == 2:  # 7.22 A edges: scale up until the smallest height clears the list cutoff
        from quansino_b200.batch import cell_heights

        f = 1.02 * calc.cutoff / cell_heights(cells).min()
        pos, cells = pos * f, cells * f
    atoms = BatchedAtoms(pos.copy(), cells, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), calc)
    batch = ReplicaBatch(atoms, calc)
    assert batch.triclinic
    e = batch.energy_full().cpu().numpy()
    e_atoms = batch.atom_energies().cpu().numpy()
    for r in range(R):
        ref = capi.energy(par, pos[r], cells[r], (1, 1, 1))
        assert abs(e[r] - ref["energy"]) <= 1e-10 * max(1.0, abs(ref["energy"])), (r, e[r], ref["energy"])
        assert abs(e_atoms[r, :n].sum() - ref["energy"]) <= 1e-10 * max(1.0, abs(ref["energy"]))
    assert int(batch.status.max().item()) == 0


def test_ragged_batch_in_general_cells(cuda_lib):
    """Replicas of different sizes in one padded batch (n_atoms 20 ... 108 of n_max 128, one of them EMPTY), each in its own sheared
    cell: energies, then displacement + anisotropic cell moves, against the oracle replica by replica."""
    from quansino_b200 import EMT, BatchedAtoms, ReplicaBatch, operations
    from quansino_b200.mc import Isobaric
    from quansino_b200.moves import CellMove, DisplacementMove

    pos108, cells, n = _sheared(3, 5, seed=15, strength=0.08)
    counts = np.array([108, 61, 20, 0, 97], dtype=np.int32)
    n_max = 128
    pos = np.zeros((5, n_max, 3))
    for r in range(5):
        pos[r, : counts[r]] = pos108[r, : counts[r]]
    atoms = BatchedAtoms(pos.copy(), cells, counts, "Cu", (True, True, True), EMT())
    batch = ReplicaBatch(atoms, EMT())
    e = batch.energy_full().cpu().numpy()
    par = emt_params("Cu")
    for r in range(5):
        ref = capi.energy(par, pos[r, : counts[r]], cells[r], (1, 1, 1))["energy"] if counts[r] else 0.0
        assert abs(e[r] - ref) <= 1e-10 * max(1.0, abs(ref)), (r, e[r], ref)
    n_attempts, p_cell = 80, 0.25
    mc = Isobaric(atoms, temperature=700.0, pressure=200.0 * BAR, seed=444, trace=True, resync_interval=0, max_cycles=n_attempts,
                  default_displacement_move=DisplacementMove(np.arange(n_max)), default_cell_move=CellMove(operations.AnisotropicDeformation(0.006)))  # fmt: skip
    mc.moves["default_cell_move"].probability = p_cell
    mc.moves["default_displacement_move"].probability = 1 - p_cell
    mc.run(1)
    tr, out = mc.last_trace, mc.download()
    for r in range(5):
        if counts[r] == 0:
            assert np.all(tr["accepted"][r][tr["move"][r] == 0] == -1)  # nothing to displace: the move fails, like the reference's None
            continue
        rep = capi.Replica(positions=pos[r].copy(), cell=cells[r], n_atoms=int(counts[r]), temperature=700.0, pressure=200.0 * BAR)
        lab = np.arange(n_max, dtype=np.int32)
        lab[counts[r] :] = -1
        specs = [
            capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.1, probability=1 - p_cell, labels=lab),
            capi.MoveSpec(capi.MOVE_CELL, capi.OP_ANISOTROPIC, step=0.006, probability=p_cell),
        ]
        ref = capi.mc_run(par, rep, specs, 444, r, 0, n_attempts)
        assert np.array_equal(tr["accepted"][r], ref["accepted"]), f"accept/reject sequence differs for replica {r}"
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(out.positions[r, : counts[r]], rep.positions[: counts[r]], rtol=0, atol=1e-10)
        np.testing.assert_allclose(out.cell[r], rep.cell, rtol=0, atol=1e-12)
    assert int(mc.batch.status.max().item()) == 0


def test_partially_periodic_general_cell(cuda_lib):
    """A sheared slab: periodic along the two in-plane vectors only."""
    from quansino_b200 import EMT, BatchedAtoms, ReplicaBatch

    pos, cells, n = _sheared(3, 3, seed=77, strength=0.1)
    cells[:, 2] *= 2.0  # vacuum along the third cell vector
    atoms = BatchedAtoms(pos.copy(), cells, np.full(3, n, dtype=np.int32), "Cu", (True, True, False), EMT())
    batch = ReplicaBatch(atoms, EMT())
    e = batch.energy_full().cpu().numpy()
    for r in range(3):
        ref = capi.energy(emt_params("Cu"), pos[r], cells[r], (1, 1, 0))
        assert abs(e[r] - ref["energy"]) <= 1e-10 * abs(ref["energy"])


@pytest.mark.parametrize("repeat,pot", [(3, "emt"), (4, "emt"), (3, "lj")])
def test_displacement_sweeps_in_general_cells_match_oracle(cuda_lib, repeat, pot):
    """Canonical sweeps (mc/canonical.py, moves/displacement.py:143-171) in fixed sheared cells, one shear per replica."""
    from quansino_b200 import EMT, BatchedAtoms, LennardJones, operations
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove

    R = 5
    pos, cells, n = _sheared(repeat, R, seed=60 + repeat, strength=0.1, wander=0.1)
    calc = EMT() if pot == "emt" else LennardJones(sigma=2.3, epsilon=0.2, rc=6.0)
    par = emt_params("Cu") if pot == "emt" else LJParams(sigma=2.3, epsilon=0.2, rc=6.0)
    atoms = BatchedAtoms(pos.copy(), cells, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), calc)
    mc = Canonical(atoms, temperature=900.0, seed=321, trace=True, resync_interval=0, max_cycles=2 * n,
                   default_displacement_move=DisplacementMove(np.arange(n), operations.Box(0.12)))  # fmt: skip
    mc.run(1)
    tr, out = mc.last_trace, mc.download()
    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cells[r], temperature=900.0)
        specs = [capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BOX, step=0.12, labels=np.arange(n, dtype=np.int32))]
        ref = capi.mc_run(par, rep, specs, 321, r, 0, 2 * n)
        assert np.array_equal(tr["accepted"][r], ref["accepted"]), f"accept/reject sequence differs for replica {r}"
        assert np.array_equal(tr["moved"][r], ref["moved"])
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(out.positions[r], rep.positions, rtol=0, atol=1e-10)
    # the running energy still agrees with a fresh evaluation
    e_run = mc.batch.energy.cpu().numpy().copy()
    e_new = mc.batch.energy_full().cpu().numpy()
    np.testing.assert_allclose(e_run, e_new, rtol=1e-11, atol=1e-10)
    assert int(mc.batch.status.max().item()) == 0


@pytest.mark.parametrize(
    "op_name,stress",
    [("Anisotropic", None), ("Shape", None), ("Anisotropic", "full"), ("Shape", "diag"), ("Isotropic", "full")],
)
def test_cell_moves_leaving_general_cells_match_oracle(cuda_lib, op_name, stress):
    """NPT / NST with the three deformation operations, 256-atom cells (no second images) and 108-atom cells (with), every attempt
    of every replica against the oracle: moves/cell.py:57-110, operations/cell.py:56-169, mc/criteria.py:143-219."""
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import Isobaric, Isotension
    from quansino_b200.moves import CellMove, DisplacementMove

    sigma = {None: None, "full": np.array([[3e-3, 1e-3, -5e-4], [1e-3, -2e-3, 7e-4], [-5e-4, 7e-4, 1e-3]]), "diag": np.diag([2e-3, -1e-3, 1e-3])}[stress]
    code = {"Anisotropic": capi.OP_ANISOTROPIC, "Shape": capi.OP_SHAPE, "Isotropic": capi.OP_ISOTROPIC}[op_name]
    for repeat, n_attempts in ((3, 120), (4, 60)):
        R, p_cell, max_value = 4, 0.3, 0.008
        pos, cells, n = _sheared(repeat, R, seed=90 + repeat, strength=0.05 if op_name != "Isotropic" else 0.1)
        atoms = BatchedAtoms(pos.copy(), cells, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), EMT())
        op = getattr(operations, f"{op_name}Deformation")(max_value)
        kw = dict(temperature=700.0, pressure=500.0 * BAR, seed=555, trace=True, resync_interval=0, max_cycles=n_attempts,
                  default_displacement_move=DisplacementMove(np.arange(n), operations.Ball(0.1)), default_cell_move=CellMove(op))  # fmt: skip
        mc = Isobaric(atoms, **kw) if sigma is None else Isotension(atoms, external_stress=sigma, **kw)
        mc.moves["default_cell_move"].probability = p_cell
        mc.moves["default_displacement_move"].probability = 1 - p_cell
        mc.run(1)
        tr, out = mc.last_trace, mc.download()
        par = emt_params("Cu")
        n_cell_acc = 0
        for r in range(R):
            rep = capi.Replica(positions=pos[r].copy(), cell=cells[r], temperature=700.0, pressure=500.0 * BAR, external_stress=sigma)
            specs = [
                capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.1, probability=1 - p_cell, labels=np.arange(n, dtype=np.int32)),
                capi.MoveSpec(capi.MOVE_CELL, code, step=max_value, probability=p_cell),
            ]
            ref = capi.mc_run(par, rep, specs, 555, r, 0, n_attempts)
            assert np.array_equal(tr["move"][r], ref["move"])
            assert np.array_equal(tr["accepted"][r], ref["accepted"]), f"accept/reject sequence differs for replica {r} ({n} atoms)"
            np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-10)
            np.testing.assert_allclose(out.positions[r], rep.positions, rtol=0, atol=1e-10)
            np.testing.assert_allclose(out.cell[r], rep.cell, rtol=0, atol=1e-12)
            n_cell_acc += int(((ref["move"] == 1) & (ref["accepted"] == 1)).sum())
        assert n_cell_acc >= 8, n_cell_acc
        assert int(mc.batch.status.max().item()) == 0


def test_shape_deformation_preserves_the_volume(cuda_lib):
    """ShapeDeformation is traceless (operations/cell.py:133-139): det expm(T) = 1, the volume never changes."""
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import Isobaric
    from quansino_b200.moves import CellMove, DisplacementMove

    pos0, cell0 = fcc_cubic("Cu", 3)
    n = len(pos0)
    R = 8
    pos = np.stack([rattle(pos0, 0.05, 10 + r) for r in range(R)])
    atoms = BatchedAtoms(pos, cell0, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    mc = Isobaric(atoms, temperature=800.0, pressure=0.0, seed=9, resync_interval=0, max_cycles=200,
                  default_displacement_move=DisplacementMove(np.arange(n), operations.Ball(0.1)),
                  default_cell_move=CellMove(operations.ShapeDeformation(0.01)))  # fmt: skip
    mc.moves["default_cell_move"].probability = 0.5
    mc.moves["default_displacement_move"].probability = 0.5
    mc.run(1)
    out = mc.download()
    assert mc.batch.triclinic
    accepted = mc.counters()[:, mc._slots["default_cell_move"], 1]
    assert accepted.sum() > 20
    off = out.cell.copy()
    off[:, [0, 1, 2], [0, 1, 2]] = 0.0
    assert np.abs(off).max() > 1e-3  # the cells did change shape
    np.testing.assert_allclose(np.abs(np.linalg.det(out.cell)), abs(np.linalg.det(cell0)), rtol=1e-12)


def test_grand_canonical_in_general_cells_matches_oracle(cuda_lib):
    """Plain exchange moves (moves/exchange.py:88-176, Translation placement `uniform(0, 1, (1, 3)) @ cell`) in sheared periodic
    cells, one shear per replica: a Lennard-Jones fluid that gains and loses atoms, every attempt against the oracle, label tables
    bit-exact."""
    from quansino_b200 import BatchedAtoms, LennardJones
    from quansino_b200.mc import GrandCanonical
    from quansino_b200.moves import DisplacementMove, ExchangeMove

    R, n0, n_max, n_attempts, mu, T = 4, 20, 64, 500, -0.5, 400.0
    rng = np.random.default_rng(3)
    cells = np.zeros((R, 3, 3))
    pos = np.zeros((R, n_max, 3))
    for r in range(R):
        cells[r] = (np.eye(3) + rng.uniform(-0.2, 0.2, (3, 3)) * (1 - np.eye(3))) @ (np.eye(3) * 12.0)
        grid = np.array([[i, j, k] for i in range(3) for j in range(3) for k in range(3)], dtype=float)[:n0]
        pos[r, :n0] = ((grid + 0.5) / 3.0 + rng.normal(scale=0.02, size=(n0, 3))) @ cells[r]
    calc = LennardJones(sigma=2.5, epsilon=0.05, rc=6.0)
    par = LJParams(sigma=2.5, epsilon=0.05, rc=6.0)
    atoms = BatchedAtoms(pos.copy(), cells, np.full(R, n0, dtype=np.int32), "Pt", (True, True, True), calc)
    lab = np.arange(n0)
    mc = GrandCanonical(
        atoms, exchange_atoms=("Pt", (0.0, 0.0, 0.0)), temperature=T, chemical_potential=mu, number_of_exchange_particles=n0,
        max_cycles=n_attempts, seed=31, trace=True, resync_interval=0,
        default_displacement_move=DisplacementMove(lab), default_exchange_move=ExchangeMove(lab),
    )  # fmt: skip
    mc.moves["default_displacement_move"].probability = 0.4
    mc.moves["default_exchange_move"].probability = 0.6
    vol = np.abs(np.linalg.det(cells))
    assert mc.batch.triclinic
    mc.accessible_volume = float(vol[0])
    mc.run(1)
    tr, out = mc.last_trace, mc.download()
    n_ex = mc.number_of_exchange_particles
    labels_d, labels_x = mc.labels("default_displacement_move"), mc.labels("default_exchange_move")
    ins = dele = 0
    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cells[r], pbc=(1, 1, 1), n_atoms=n0, temperature=T, mass=195.084, chemical_potential=mu,
                           exchange_mass=195.084, n_exchange=n0, accessible_volume=float(vol[0]))  # fmt: skip
        ld = np.full(n_max, -1, np.int32)
        ld[:n0] = lab
        lx = ld.copy()
        specs = [
            capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.1, probability=0.4, labels=ld),
            capi.MoveSpec(capi.MOVE_EXCHANGE, capi.OP_TRANSLATION, probability=0.6, labels=lx),
        ]
        ref = capi.mc_run(par, rep, specs, 31, r, 0, n_attempts)
        assert np.array_equal(tr["move"][r], ref["move"])
        assert np.array_equal(tr["accepted"][r], ref["accepted"]), f"accept/reject sequence differs for replica {r}"
        assert np.array_equal(tr["moved"][r], ref["moved"])
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-10)
        assert out.n_atoms[r] == rep.n_atoms and n_ex[r] == rep.n_exchange
        assert np.array_equal(labels_d[r, : rep.n_atoms], ld[: rep.n_atoms])
        assert np.array_equal(labels_x[r, : rep.n_atoms], lx[: rep.n_atoms])
        np.testing.assert_allclose(out.positions[r, : rep.n_atoms], rep.positions[: rep.n_atoms], rtol=0, atol=1e-10)
        d_n = np.diff(np.concatenate([[n0], ref["n_atoms"]]))
        ins += int((d_n == 1).sum())
        dele += int((d_n == -1).sum())
    assert ins > 20 and dele > 20, (ins, dele)
    assert int(mc.batch.status.max().item()) == 0


@pytest.mark.parametrize("build_lanes", [1, 4, 32])
@pytest.mark.parametrize("repeat", [3, 4])
def test_forces_in_general_cells_match_oracle(cuda_lib, monkeypatch, build_lanes, repeat):
    """The Verlet-list force kernels (csrc/hmc_kernel.cuh) in sheared cells: list shifts count cell vectors, the build rounds in
    fractional coordinates; 108 atoms (several images per pair) and 256 atoms, atoms that wandered by whole cell vectors."""
    from quansino_b200 import EMT, BatchedAtoms, ReplicaBatch

    monkeypatch.setenv("QMC_HMC_BUILD_LANES", str(build_lanes))
    R = 3
    pos, cells, n = _sheared(repeat, R, seed=70 + repeat, strength=0.1, wander=0.1)
    batch = ReplicaBatch(BatchedAtoms(pos.copy(), cells, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), EMT()), EMT())
    f = batch.forces_full().cpu().numpy().transpose(0, 2, 1)
    e = batch.energy.cpu().numpy()
    for r in range(R):
        ref = capi.energy(emt_params("Cu"), pos[r], cells[r], (1, 1, 1), want_forces=True)
        np.testing.assert_allclose(f[r], ref["forces"], rtol=1e-9, atol=1e-10)
        assert abs(e[r] - ref["energy"]) <= 1e-10 * abs(ref["energy"])
    assert int(batch.status.max().item()) == 0


def test_large_replica_in_a_general_cell(cuda_lib):
    """2048 atoms (beyond the shared-memory kernels: energies come from the Verlet-list kernels) in a sheared cell: energy and forces
    against the oracle's brute-force image enumeration."""
    from quansino_b200 import EMT, BatchedAtoms, ReplicaBatch

    pos, cells, n = _sheared(8, 1, seed=88, strength=0.05, stdev=0.03)
    assert n == 2048
    batch = ReplicaBatch(BatchedAtoms(pos.copy(), cells, np.full(1, n, dtype=np.int32), "Cu", (True, True, True), EMT()), EMT())
    assert batch.triclinic
    e = batch.energy_full().cpu().numpy()
    f = batch.forces_full().cpu().numpy().transpose(0, 2, 1)
    ref = capi.energy(emt_params("Cu"), pos[0], cells[0], (1, 1, 1), want_forces=True)
    assert abs(e[0] - ref["energy"]) <= 1e-10 * abs(ref["energy"])
    np.testing.assert_allclose(f[0], ref["forces"], rtol=1e-9, atol=1e-10)
    assert int(batch.status.max().item()) == 0


def test_hmc_and_force_bias_in_general_cells_match_oracle(cuda_lib):
    """Hamiltonian moves (integrators/displacement.py:52-81, mc/criteria.py:117-140) and ForceBias steps (mc/fbmc.py:202-251) in
    sheared cells, one shear per replica, against the oracle."""
    from quansino_b200 import EMT, BatchedAtoms
    from quansino_b200.integrators import Verlet
    from quansino_b200.integrators.displacement import fs
    from quansino_b200.mc import ForceBias, HamiltonianCanonical
    from quansino_b200.moves import HamiltonianDisplacementMove

    R, n_attempts, n_steps = 4, 3, 12
    pos, cells, n = _sheared(3, R, seed=81, strength=0.08, stdev=0.03)
    temps = np.array([300.0, 500.0, 800.0, 1200.0])
    atoms = BatchedAtoms(pos.copy(), cells, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    mc = HamiltonianCanonical(
        atoms, temperature=temps, max_cycles=1, seed=98, trace=True, resync_interval=0, steps_per_launch=n_attempts,
        default_displacement_move=HamiltonianDisplacementMove(operation=Verlet(dt=2.0, max_steps=n_steps)),
    )  # fmt: skip
    mc.run(n_attempts)
    tr, out, par = mc.last_trace, mc.download(), emt_params("Cu")
    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cells[r], temperature=float(temps[r]), mass=63.546, momenta=np.zeros((n, 3)))
        specs = [capi.MoveSpec(capi.MOVE_HMC, capi.OP_VERLET, dt=2.0 * fs, n_steps=n_steps)]
        ref = capi.mc_run(par, rep, specs, 98, r, 0, n_attempts)
        assert np.array_equal(tr["accepted"][r], ref["accepted"])
        np.testing.assert_allclose(tr["delta"][r], ref["delta"], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(out.positions[r], rep.positions, rtol=0, atol=1e-10)
        np.testing.assert_allclose(out.momenta[r], rep.momenta, rtol=0, atol=1e-10)
    assert tr["accepted"].sum() > 0
    steps = 3
    fb = ForceBias(BatchedAtoms(pos.copy(), cells, np.full(R, n, np.int32), "Cu", (True,) * 3, EMT()), delta=0.08, temperature=temps, seed=76)
    fb.run(steps)
    out, e_gpu = fb.download(), fb.context.last_potential_energy
    for r in range(R):
        p = pos[r].copy()
        for s in range(steps):
            e, _, _ = capi.fbmc_step(par, p, cells[r], (1, 1, 1), 63.546, float(temps[r]), 0.08, 76, r, s)
        np.testing.assert_allclose(out.positions[r], p, rtol=0, atol=1e-11)
        assert abs(e_gpu[r] - e) <= 1e-10 * abs(e)


def test_general_cells_refuse_what_their_kernels_do_not_hold(cuda_lib):
    from quansino_b200 import EMT, BatchedAtoms, ReplicaBatch
    from quansino_b200.mc import GrandCanonical
    from quansino_b200.moves import DisplacementMove, ExchangeMove

    pos, cells, n = _sheared(3, 2, seed=5)
    pos_pad = np.zeros((2, 128, 3))
    pos_pad[:, :n] = pos
    atoms = BatchedAtoms(pos_pad, cells, np.full(2, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    dimer = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 2.5]])  # a molecular species: run-coded label groups, not in the general-cell kernels
    gc = GrandCanonical(atoms, exchange_atoms=("Cu", dimer), temperature=1000.0, chemical_potential=-3.0, number_of_exchange_particles=0, max_cycles=10,
                        seed=1, default_displacement_move=DisplacementMove(np.arange(n)), default_exchange_move=ExchangeMove(np.full(n, -1)))  # fmt: skip
    with pytest.raises(NotImplementedError, match="triclinic"):
        gc.run(1)
    small = cells.copy()
    small[:, 0] *= 0.5  # a height below the cutoff
    with pytest.raises(NotImplementedError, match="height"):
        ReplicaBatch(BatchedAtoms(pos.copy(), small, np.full(2, n, dtype=np.int32), "Cu", (True, True, True), EMT()), EMT())
