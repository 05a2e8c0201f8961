"""Parity of the isobaric, grand-canonical and Hamiltonian paths against the CPU oracle (configs C3 / C4 / C5 of
BASELINE.json at test sizes), plus the parallel-tempering swap kernel.

Bar (north_star): energies within 1e-10 relative, identical accept / reject sequences, bit-exact integer state."""

import numpy as np
import pytest

from oracle import capi
from oracle.potentials import LJParams, emt_params
from tests.helpers import cu_replicas, oracle_par, oracle_run

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_variant")]

BAR = 6.241509125883258e-07  # ase.units.bar


def _emt_atoms(pos, cell, n):
    from quansino_b200 import EMT, BatchedAtoms

    return BatchedAtoms(pos.copy(), cell, np.full(len(pos), n, dtype=np.int32), "Cu", (True, True, True), EMT())


@pytest.mark.parametrize("p_cell", [None, 0.15])
def test_isobaric_matches_oracle(cuda_lib, p_cell):
    from quansino_b200.mc import Isobaric
    from quansino_b200.moves import CellMove, DisplacementMove
    from quansino_b200.operations import Ball, IsotropicDeformation

    R, steps = 6, 2
    pos, cell, n = cu_replicas(3, R, seed=11)
    pressure = 1000.0 * BAR
    mc = Isobaric(
        _emt_atoms(pos, cell, n), temperature=300.0, pressure=pressure, seed=77, trace=True, steps_per_launch=steps, resync_interval=0,
        default_displacement_move=DisplacementMove(np.arange(n), Ball(0.1)), default_cell_move=CellMove(IsotropicDeformation(0.05)),
    )  # fmt: skip
    if p_cell is None:
        assert mc.moves["default_cell_move"].probability == 1 / (n + 1)  # mc/isobaric.py:124-127
        assert mc.moves["default_displacement_move"].probability == 1 / (1 + 1 / n)
    else:
        mc.moves["default_cell_move"].probability = p_cell
        mc.moves["default_displacement_move"].probability = 1 - p_cell
    probs = [mc.moves["default_displacement_move"].probability, mc.moves["default_cell_move"].probability]
    mc.run(steps)
    tr = mc.last_trace
    out = mc.download()
    par = emt_params("Cu")
    n_cell_moves = 0
    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cell, temperature=300.0, pressure=pressure)
        specs = [
            capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.1, probability=probs[0], labels=np.arange(n, dtype=np.int32)),
            capi.MoveSpec(capi.MOVE_CELL, capi.OP_ISOTROPIC, step=0.05, probability=probs[1]),
        ]
        ref = capi.mc_run(par, rep, specs, 77, r, 0, steps * n)
        n_cell_moves += int((ref["move"] == 1).sum())
        assert np.array_equal(tr["move"][r], ref["move"])
        assert np.array_equal(tr["accepted"][r], ref["accepted"]), f"accept/reject sequence differs for replica {r}"
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(out.positions[r], rep.positions, rtol=0, atol=1e-11)
        np.testing.assert_allclose(out.cell[r], rep.cell, rtol=1e-14)
    if p_cell is not None:
        assert n_cell_moves > 20  # the cell branch really ran


def _gc_setup(R, n_max, mu):
    from quansino_b200 import BatchedAtoms, LennardJones
    from quansino_b200.systems import octahedron, rattle

    pos0, cell = octahedron("Pt", 4, 10.0)
    n = len(pos0)
    pos = np.zeros((R, n_max, 3))
    for r in range(R):
        pos[r, :n] = rattle(pos0, 0.02, 500 + r)
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), "Pt", (False, False, False), LennardJones(sigma=2.5, epsilon=0.1))
    return atoms, pos, cell, n


def test_grand_canonical_lj_matches_oracle(cuda_lib):
    from quansino_b200.mc import GrandCanonical
    from quansino_b200.moves import DisplacementMove, ExchangeMove

    R, n_max, mu, n_attempts = 6, 64, -0.5, 1500
    atoms, pos, cell, n = _gc_setup(R, n_max, mu)
    lab_d = np.arange(n)
    lab_x = np.full(n, -1)
    mc = GrandCanonical(
        atoms, exchange_atoms=("Pt", (0.0, 0.0, 0.0)), temperature=300.0, chemical_potential=mu, number_of_exchange_particles=0,
        max_cycles=n_attempts, seed=11, trace=True, resync_interval=0,
        default_displacement_move=DisplacementMove(lab_d), default_exchange_move=ExchangeMove(lab_x),
    )  # fmt: skip
    mc.moves["default_displacement_move"].probability = 0.5
    mc.moves["default_exchange_move"].probability = 0.5
    mc.run(1)
    tr = mc.last_trace
    out = mc.download()
    n_ex = mc.number_of_exchange_particles
    labels_d, labels_x = mc.labels("default_displacement_move"), mc.labels("default_exchange_move")
    par = LJParams(sigma=2.5, epsilon=0.1)
    volume = float(np.prod(np.diag(cell)))
    seen = {"insert": 0, "delete": 0, "failed": 0}
    for r in range(R):
        P = pos[r].copy()
        rep = capi.Replica(positions=P, cell=cell, pbc=(0, 0, 0), n_atoms=n, temperature=300.0, mass=195.084, chemical_potential=mu,
                           exchange_mass=195.084, n_exchange=0, accessible_volume=volume)  # fmt: skip
        ld = np.full(n_max, -1, np.int32)
        ld[:n] = lab_d
        lx = np.full(n_max, -1, np.int32)
        specs = [
            capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.1, probability=0.5, labels=ld),
            capi.MoveSpec(capi.MOVE_EXCHANGE, capi.OP_TRANSLATION, probability=0.5, labels=lx),
        ]
        ref = capi.mc_run(par, rep, specs, 11, r, 0, n_attempts)
        assert np.array_equal(tr["move"][r], ref["move"])
        assert np.array_equal(tr["accepted"][r], ref["accepted"]), f"accept/reject sequence differs for replica {r}"
        assert np.array_equal(tr["moved"][r], ref["moved"])  # bit-exact indices
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-11)
        # bit-exact integer state: atom count, exchange count, both label tables
        assert out.n_atoms[r] == rep.n_atoms
        assert n_ex[r] == rep.n_exchange
        assert np.array_equal(labels_d[r, : rep.n_atoms], ld[: rep.n_atoms])
        assert np.array_equal(labels_x[r, : rep.n_atoms], lx[: rep.n_atoms])
        np.testing.assert_allclose(out.positions[r, : rep.n_atoms], rep.positions[: rep.n_atoms], rtol=0, atol=1e-11)
        ex = ref["move"] == 1
        d_n = np.diff(np.concatenate([[n], ref["n_atoms"]]))
        seen["insert"] += int((d_n == 1).sum())
        seen["delete"] += int((d_n == -1).sum())
        seen["failed"] += int((ref["accepted"][ex] == -1).sum())
    assert seen["insert"] > 10 and seen["delete"] > 5 and seen["failed"] > 5, seen


def test_grand_canonical_emt_matches_oracle(cuda_lib):
    """Insert / delete with the many-body potential: periodic Cu with one vacancy-rich cell, EMT."""
    from quansino_b200 import EMT, BatchedAtoms
    from quansino_b200.mc import GrandCanonical
    from quansino_b200.moves import DisplacementMove, ExchangeMove

    R, n_max, n_attempts, mu = 4, 128, 400, -3.4
    pos108, cell, n = cu_replicas(3, R, seed=5, n_max=n_max)
    atoms = BatchedAtoms(pos108.copy(), cell, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    lab = np.arange(n)
    mc = GrandCanonical(
        atoms, exchange_atoms=("Cu", (0.0, 0.0, 0.0)), temperature=3000.0, chemical_potential=mu, number_of_exchange_particles=n,
        max_cycles=n_attempts, seed=5, trace=True, resync_interval=0,
        default_displacement_move=DisplacementMove(lab), default_exchange_move=ExchangeMove(lab),
    )  # fmt: skip
    mc.moves["default_displacement_move"].probability = 0.3
    mc.moves["default_exchange_move"].probability = 0.7
    mc.run(1)
    tr = mc.last_trace
    out = mc.download()
    par = emt_params("Cu")
    volume = float(np.prod(np.diag(cell)))
    changed = 0
    for r in range(R):
        rep = capi.Replica(positions=pos108[r].copy(), cell=cell, n_atoms=n, temperature=3000.0, mass=63.546, chemical_potential=mu,
                           exchange_mass=63.546, n_exchange=n, accessible_volume=volume)  # fmt: skip
        ld = np.full(n_max, -1, np.int32)
        ld[:n] = lab
        lx = ld.copy()
        specs = [
            capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.1, probability=0.3, labels=ld),
            capi.MoveSpec(capi.MOVE_EXCHANGE, capi.OP_TRANSLATION, probability=0.7, labels=lx),
        ]
        ref = capi.mc_run(par, rep, specs, 5, r, 0, n_attempts)
        assert np.array_equal(tr["accepted"][r], ref["accepted"])
        assert np.array_equal(tr["moved"][r], ref["moved"])
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-10)
        assert out.n_atoms[r] == rep.n_atoms
        changed += int((np.diff(ref["n_atoms"]) != 0).sum())
        np.testing.assert_allclose(out.positions[r, : rep.n_atoms], rep.positions[: rep.n_atoms], rtol=0, atol=1e-11)
    assert changed > 5


def test_molecular_exchange_matches_oracle(cuda_lib):
    """Molecular exchange on the device (moves/exchange.py:88-176 with a three-atom `exchange_atoms`, placed by
    TranslationRotation as the docstring at :194-203 prescribes; reference tests tests/mc/test_grand_canonical.py:359-440): a Cu
    cluster in a periodic box with the many-body potential, trimers inserted and deleted, displacement moves that see the inserted
    molecules as label groups -- traces, atom counts, particle counts and BOTH label tables bit-exact against the C oracle."""
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import GrandCanonical
    from quansino_b200.moves import DisplacementMove, ExchangeMove
    from quansino_b200.systems import fcc_cubic, rattle

    R, n_max, n_attempts, mu, L = 4, 96, 300, -0.5, 16.0
    pos0, cell0 = fcc_cubic("Cu", 2)
    n = len(pos0)
    cell = np.eye(3) * L
    pos0 = pos0 + (L - cell0[0, 0]) / 2
    pos = np.zeros((R, n_max, 3))
    for r in range(R):
        pos[r, :n] = rattle(pos0, 0.03, 700 + r)
    tri = np.array([[0.0, 0.0, 0.0], [2.4, 0.0, 0.0], [1.2, 2.0, 0.3]])
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    mc = GrandCanonical(
        atoms, exchange_atoms=("Cu", tri), temperature=2500.0, chemical_potential=mu, number_of_exchange_particles=0,
        max_cycles=n_attempts, seed=21, trace=True, resync_interval=0,
        default_displacement_move=DisplacementMove(np.arange(n), operations.Ball(0.15)),
        default_exchange_move=ExchangeMove(np.full(n, -1), operations.TranslationRotation()),
    )  # fmt: skip
    mc.accessible_volume = L**3
    mc.moves["default_displacement_move"].probability = 0.4
    mc.moves["default_exchange_move"].probability = 0.6
    mc.run(1)
    tr, out = mc.last_trace, mc.download()
    n_ex = mc.number_of_exchange_particles
    labels_d, labels_x = mc.labels("default_displacement_move"), mc.labels("default_exchange_move")
    par = emt_params("Cu")
    inserted = deleted = group_moves = 0
    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cell, pbc=(1, 1, 1), n_atoms=n, temperature=2500.0, mass=63.546, chemical_potential=mu,
                           exchange_mass=3 * 63.546, n_exchange=0, accessible_volume=L**3, exchange_mol=tri)  # fmt: skip
        ld = np.full(n_max, -1, np.int32)
        ld[:n] = np.arange(n)
        lx = np.full(n_max, -1, np.int32)
        specs = [
            capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.15, probability=0.4, labels=ld),
            capi.MoveSpec(capi.MOVE_EXCHANGE, capi.OP_TRANSLATION_ROTATION, probability=0.6, labels=lx),
        ]
        ref = capi.mc_run(par, rep, specs, 21, r, 0, n_attempts)
        assert np.array_equal(tr["move"][r], ref["move"])
        assert np.array_equal(tr["accepted"][r], ref["accepted"]), f"accept/reject sequence differs for replica {r}"
        assert np.array_equal(tr["moved"][r], ref["moved"])
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-10)
        assert out.n_atoms[r] == rep.n_atoms and n_ex[r] == rep.n_exchange and (rep.n_atoms - n) == 3 * rep.n_exchange
        assert np.array_equal(labels_d[r, : rep.n_atoms], ld[: rep.n_atoms])
        assert np.array_equal(labels_x[r, : rep.n_atoms], lx[: rep.n_atoms])
        np.testing.assert_allclose(out.positions[r, : rep.n_atoms], rep.positions[: rep.n_atoms], rtol=0, atol=1e-10)
        d_n = np.diff(np.concatenate([[n], ref["n_atoms"]]))
        inserted += int((d_n == 3).sum())
        deleted += int((d_n == -3).sum())
        group_moves += int(((ref["move"] == 0) & (ref["moved"] >= n) & (ref["accepted"] == 1)).sum())
    assert inserted > 20 and deleted > 20 and group_moves > 0, (inserted, deleted, group_moves)
    assert int(mc.batch.status.max().item()) == 0


def test_composite_exchange_matches_oracle(cuda_lib):
    """`CompositeExchangeMove` on the device (moves/exchange.py:237-316; reference test tests/moves/test_composite_moves.py:264-296):
    three single-atom exchange moves under one coin and ONE criteria evaluation with particle_delta = +-(sub-moves that acted),
    Cu cluster in a periodic box with the many-body potential.  An accepted insertion gives its three atoms ONE label, so later
    deletions take out three atoms per label (one, two or three labels at a time) and displacement moves translate them together --
    traces, atom counts, particle counts and both label tables bit-exact against the C oracle."""
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import GrandCanonical
    from quansino_b200.mc.criteria import GrandCanonicalCriteria
    from quansino_b200.moves import CompositeExchangeMove, DisplacementMove, ExchangeMove
    from quansino_b200.systems import fcc_cubic, rattle

    R, n_max, n_attempts, mu, L, k = 4, 96, 300, -0.5, 16.0, 3
    pos0, cell0 = fcc_cubic("Cu", 2)
    n = len(pos0)
    cell = np.eye(3) * L
    pos0 = pos0 + (L - cell0[0, 0]) / 2
    pos = np.zeros((R, n_max, 3))
    for r in range(R):
        pos[r, :n] = rattle(pos0, 0.03, 800 + r)
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    exch = ExchangeMove(np.full(n, -1)) + ExchangeMove(np.full(n, -1)) + ExchangeMove(np.full(n, -1))
    assert isinstance(exch, CompositeExchangeMove) and len(exch) == k
    mc = GrandCanonical(
        atoms, exchange_atoms="Cu", temperature=2500.0, chemical_potential=mu, number_of_exchange_particles=0,
        max_cycles=n_attempts, seed=22, trace=True, resync_interval=0,
        default_displacement_move=DisplacementMove(np.arange(n), operations.Ball(0.15)),
    )  # fmt: skip
    mc.add_move(exch, criteria=GrandCanonicalCriteria(), name="default_exchange_move")
    mc.accessible_volume = L**3
    mc.moves["default_displacement_move"].probability = 0.4
    mc.moves["default_exchange_move"].probability = 0.6
    mc.run(1)
    tr, out = mc.last_trace, mc.download()
    n_ex = mc.number_of_exchange_particles
    labels_d, labels_x = mc.labels("default_displacement_move"), mc.labels("default_exchange_move")
    par = emt_params("Cu")
    changes: dict[int, int] = {}
    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cell, pbc=(1, 1, 1), n_atoms=n, temperature=2500.0, mass=63.546, chemical_potential=mu,
                           exchange_mass=63.546, n_exchange=0, accessible_volume=L**3)  # fmt: skip
        ld = np.full(n_max, -1, np.int32)
        ld[:n] = np.arange(n)
        lx = np.full(n_max, -1, np.int32)
        specs = [
            capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.15, probability=0.4, labels=ld),
            capi.MoveSpec(capi.MOVE_EXCHANGE, capi.OP_TRANSLATION, probability=0.6, labels=lx, repeat=k, bias_insert=0.5),
        ]
        ref = capi.mc_run(par, rep, specs, 22, r, 0, n_attempts)
        assert np.array_equal(tr["move"][r], ref["move"])
        assert np.array_equal(tr["accepted"][r], ref["accepted"]), f"accept/reject sequence differs for replica {r}"
        assert np.array_equal(tr["moved"][r], ref["moved"])
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-10)
        assert out.n_atoms[r] == rep.n_atoms and n_ex[r] == rep.n_exchange
        assert np.array_equal(labels_d[r, : rep.n_atoms], ld[: rep.n_atoms])
        assert np.array_equal(labels_x[r, : rep.n_atoms], lx[: rep.n_atoms])
        np.testing.assert_allclose(out.positions[r, : rep.n_atoms], rep.positions[: rep.n_atoms], rtol=0, atol=1e-10)
        d_n = np.diff(np.concatenate([[n], ref["n_atoms"]]))
        for v, c in zip(*np.unique(d_n, return_counts=True)):
            changes[int(v)] = changes.get(int(v), 0) + int(c)
    assert changes.get(3, 0) > 20 and changes.get(-3, 0) > 10 and changes.get(-6, 0) > 0, changes
    assert int(mc.batch.status.max().item()) == 0

@pytest.mark.parametrize("build_lanes", [1, 4, 32])
def test_forces_match_oracle(cuda_lib, monkeypatch, build_lanes):
    """``build_lanes``: lanes per atom of the neighbour-list build (hmc_build_kernel / hmc_build_split_kernel); the library
    reads QMC_HMC_BUILD_LANES at every call."""
    from quansino_b200 import ReplicaBatch

    monkeypatch.setenv("QMC_HMC_BUILD_LANES", str(build_lanes))
    R = 4
    pos, cell, n = cu_replicas(3, R, seed=21)
    batch = ReplicaBatch(_emt_atoms(pos, cell, n), temperature=300.0)
    f = batch.forces_full().cpu().numpy().transpose(0, 2, 1)
    e = batch.energy.cpu().numpy()
    par = emt_params("Cu")
    for r in range(R):
        ref = capi.energy(par, pos[r], cell, (1, 1, 1), want_forces=True)
        np.testing.assert_allclose(f[r], ref["forces"], rtol=1e-9, atol=1e-11)
        assert abs(e[r] - ref["energy"]) <= 1e-10 * abs(ref["energy"])


def test_neighbour_list_build_variants_agree(cuda_lib, monkeypatch):
    """500-atom cell (nearest images only): the list built by 1, 8 or 32 lanes per atom has the same entries in the same order, so
    forces and energies are bit-identical; another skin lists other entries, the physics stays (1e-12)."""
    from quansino_b200 import ReplicaBatch

    pos, cell, n = cu_replicas(5, 2, seed=33)
    out = {}
    for lanes, skin in ((1, "0.7"), (8, "0.7"), (32, "0.7"), (32, "0.25"), (1, "1.3")):
        monkeypatch.setenv("QMC_HMC_BUILD_LANES", str(lanes))
        monkeypatch.setenv("QMC_HMC_SKIN", skin)
        batch = ReplicaBatch(_emt_atoms(pos, cell, n), temperature=300.0)
        out[lanes, skin] = (batch.forces_full().cpu().numpy(), batch.energy.cpu().numpy().copy())
    f0, e0 = out[1, "0.7"]
    assert np.abs(f0).max() > 0.01
    for key in ((8, "0.7"), (32, "0.7")):
        assert np.array_equal(out[key][0], f0) and np.array_equal(out[key][1], e0)
    for key in ((32, "0.25"), (1, "1.3")):
        np.testing.assert_allclose(out[key][0], f0, rtol=0, atol=1e-12)
        np.testing.assert_allclose(out[key][1], e0, rtol=1e-13)


def test_hmc_matches_oracle(cuda_lib):
    from quansino_b200.integrators import Verlet
    from quansino_b200.mc import HamiltonianCanonical
    from quansino_b200.moves import HamiltonianDisplacementMove

    R, n_attempts, n_steps = 4, 3, 12
    pos, cell, n = cu_replicas(3, R, seed=31, stdev=0.03)
    temps = np.array([300.0, 500.0, 800.0, 1200.0])
    mc = HamiltonianCanonical(
        _emt_atoms(pos, cell, n), temperature=temps, max_cycles=1, seed=99, trace=True, resync_interval=0, steps_per_launch=n_attempts,
        default_displacement_move=HamiltonianDisplacementMove(operation=Verlet(dt=2.0, max_steps=n_steps)),
    )  # fmt: skip
    mc.run(n_attempts)
    tr = mc.last_trace
    out = mc.download()
    par = emt_params("Cu")
    from quansino_b200.integrators.displacement import fs

    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cell, temperature=float(temps[r]), mass=63.546, momenta=np.zeros((n, 3)))
        specs = [capi.MoveSpec(capi.MOVE_HMC, capi.OP_VERLET, dt=2.0 * fs, n_steps=n_steps)]
        ref = capi.mc_run(par, rep, specs, 99, r, 0, n_attempts)
        assert np.array_equal(tr["accepted"][r], ref["accepted"])
        np.testing.assert_allclose(tr["delta"][r], ref["delta"], rtol=1e-9, atol=1e-10)  # differences of sums of O(100 eV) terms
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(out.positions[r], rep.positions, rtol=0, atol=1e-10)
        np.testing.assert_allclose(out.momenta[r], rep.momenta, rtol=0, atol=1e-10)
        assert abs(mc.batch.kinetic.cpu().numpy()[r] - rep.last_kinetic) <= 1e-9 * abs(rep.last_kinetic)
    assert tr["accepted"].sum() > 0


def test_hmc_c5_shape_matches_oracle(cuda_lib, kernel_variant):
    """BASELINE.json configuration C5 at its real shape: ONE replica of 2048-atom Cu (8 x 8 x 8), Verlet 100 x 1 fs, two
    trajectories -- accept sequence identical, energies / positions / momenta to 1e-10 against the C oracle
    (integrators/displacement.py:52-81, mc/criteria.py:117-140).  The oracle needs ~1 min for its 202 force evaluations."""
    from quansino_b200.integrators import Verlet
    from quansino_b200.integrators.displacement import fs
    from quansino_b200.mc import HamiltonianCanonical
    from quansino_b200.moves import HamiltonianDisplacementMove

    if kernel_variant != "warp":
        pytest.skip("the Hamiltonian kernels do not depend on the sweep-kernel variant; run once")
    n_attempts, n_steps = 2, 100
    pos, cell, n = cu_replicas(8, 1, seed=51, stdev=0.03)
    assert n == 2048
    mc = HamiltonianCanonical(
        _emt_atoms(pos, cell, n), temperature=600.0, max_cycles=1, seed=2024, trace=True, resync_interval=0, steps_per_launch=n_attempts,
        default_displacement_move=HamiltonianDisplacementMove(operation=Verlet(dt=1.0, max_steps=n_steps)),
    )  # fmt: skip
    mc.run(n_attempts)
    tr, out = mc.last_trace, mc.download()
    rep = capi.Replica(positions=pos[0].copy(), cell=cell, temperature=600.0, mass=63.546, momenta=np.zeros((n, 3)))
    specs = [capi.MoveSpec(capi.MOVE_HMC, capi.OP_VERLET, dt=1.0 * fs, n_steps=n_steps)]
    ref = capi.mc_run(emt_params("Cu"), rep, specs, 2024, 0, 0, n_attempts)
    assert np.array_equal(tr["accepted"][0], ref["accepted"])
    np.testing.assert_allclose(tr["delta"][0], ref["delta"], rtol=1e-9, atol=1e-9)  # H = O(1e3) eV: 1e-9 absolute is 1e-12 of H
    np.testing.assert_allclose(tr["energy"][0], ref["energy"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(out.positions[0], rep.positions, rtol=0, atol=1e-10)
    np.testing.assert_allclose(out.momenta[0], rep.momenta, rtol=0, atol=1e-10)
    assert abs(mc.batch.kinetic.cpu().numpy()[0] - rep.last_kinetic) <= 1e-10 * abs(rep.last_kinetic)
    assert int(mc.batch.status.max().item()) == 0


def test_parallel_tempering_round_matches_oracle_replay(cuda_lib, kernel_variant):
    """The whole parallel-tempering round of C5 on one device -- HMC trajectory of every replica at its CURRENT temperature, then
    `ReplicaExchange.exchange` (all-gather is the identity here) -- for 8 temperatures and 5 rounds against a replay with the
    C oracle: trajectories by `qo_mc_run`, swap decisions by `qo_pt_swap_accept` from the shared stream."""
    from oracle import capi as oc
    from quansino_b200.integrators import Verlet
    from quansino_b200.integrators.displacement import fs
    from quansino_b200.mc import HamiltonianCanonical
    from quansino_b200.mc.tempering import ReplicaExchange
    from quansino_b200.moves import HamiltonianDisplacementMove

    if kernel_variant != "warp":
        pytest.skip("the Hamiltonian kernels do not depend on the sweep-kernel variant; run once")
    R, rounds, n_steps, seed = 8, 5, 12, 77
    pos, cell, n = cu_replicas(3, R, seed=61, stdev=0.03)
    temps0 = 300.0 * (400.0 / 300.0) ** (np.arange(R) / (R - 1))
    mc = HamiltonianCanonical(
        _emt_atoms(pos, cell, n), temperature=temps0, max_cycles=1, seed=seed, resync_interval=0,
        default_displacement_move=HamiltonianDisplacementMove(operation=Verlet(dt=2.0, max_steps=n_steps)),
    )  # fmt: skip
    pt = ReplicaExchange.for_driver(mc)
    par = emt_params("Cu")
    reps = [capi.Replica(positions=pos[r].copy(), cell=cell, temperature=float(temps0[r]), mass=63.546, momenta=np.zeros((n, 3))) for r in range(R)]
    specs = [capi.MoveSpec(capi.MOVE_HMC, capi.OP_VERLET, dt=2.0 * fs, n_steps=n_steps)]
    expect = temps0.copy()
    n_swaps = 0
    for rnd in range(rounds):
        mc.run(1)
        pt.exchange()
        for r in range(R):  # the replay: same trajectory, same temperature, same draws
            reps[r].temperature = float(expect[r])
            capi.mc_run(par, reps[r], specs, seed, r, rnd, 1)
        energies = np.array([rep.last_energy for rep in reps])
        np.testing.assert_allclose(mc.batch.energy.cpu().numpy(), energies, rtol=1e-10, atol=1e-10)
        order = np.argsort(expect, kind="stable")
        for k in range(rnd & 1, R - 1, 2):
            ra, rb = order[k], order[k + 1]
            u = oc.uniform_pair(seed, rnd, k, 3)[1]
            if oc.lib().qo_pt_swap_accept(energies[ra], energies[rb], expect[ra], expect[rb], u):
                expect[ra], expect[rb] = expect[rb], expect[ra]
                n_swaps += 1
        assert np.array_equal(mc.batch.temperature.cpu().numpy(), expect), f"ladder after round {rnd}"
    out = mc.download()
    for r in range(R):
        np.testing.assert_allclose(out.positions[r], reps[r].positions, rtol=0, atol=1e-10)
    assert n_swaps >= 3 and pt.n_accepted == n_swaps
    assert np.array_equal(np.sort(expect), np.sort(temps0))


def test_parallel_tempering_swap_matches_oracle(cuda_lib):
    import ctypes as C

    import torch

    from oracle import capi as oc
    from quansino_b200 import _lib

    lib = _lib.load()
    n = 64
    rng = np.random.default_rng(3)
    temps0 = 300.0 * (1200.0 / 300.0) ** (np.arange(n) / (n - 1))
    perm = rng.permutation(n)
    temps = temps0[perm].copy()  # replica r currently holds temperature temps[r]
    energies = 2.0 + 0.004 * temps + rng.normal(scale=0.3, size=n)
    t_dev = torch.from_numpy(temps.copy()).cuda()
    e_dev = torch.from_numpy(energies).cuda()
    acc = torch.zeros(n, dtype=torch.int8, device="cuda")
    expect = temps.copy()
    n_swaps = 0
    for rnd in range(6):
        _lib.check(lib.qmc_pt_swap(t_dev.data_ptr(), e_dev.data_ptr(), n, 4242, rnd, acc.data_ptr(), None))
        order = np.argsort(expect, kind="stable")
        for k in range(rnd & 1, n - 1, 2):
            ra, rb = order[k], order[k + 1]
            u = oc.uniform_pair(4242, rnd, k, 3)[1]
            if oc.lib().qo_pt_swap_accept(energies[ra], energies[rb], expect[ra], expect[rb], u):
                expect[ra], expect[rb] = expect[rb], expect[ra]
                n_swaps += 1
        assert np.array_equal(t_dev.cpu().numpy(), expect), f"round {rnd}"
    assert n_swaps > 10
    assert np.array_equal(np.sort(expect), np.sort(temps))  # temperatures are only permuted


def test_composite_displacement_matches_oracle(cuda_lib):
    """CompositeDisplacementMove: `a * 2 + b` next to a plain move, 8 replicas, against the C oracle."""
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import Canonical
    from quansino_b200.mc.criteria import CanonicalCriteria
    from quansino_b200.moves import DisplacementMove

    R, n_attempts = 8, 150
    pos, cell, n = cu_replicas(3, R)
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    lab = np.arange(n)
    mc = Canonical(atoms, temperature=600.0, seed=99, trace=True, max_cycles=n_attempts, resync_interval=0,
                   default_displacement_move=DisplacementMove(lab, operations.Ball(0.1)))  # fmt: skip
    combo = DisplacementMove(lab, operations.Ball(0.07)) * 2 + DisplacementMove(lab, operations.Box(0.03))
    mc.add_move(combo, criteria=CanonicalCriteria(), name="combo", probability=2.0)
    mc.run(1)
    tr = mc.last_trace
    par = oracle_par(mc.batch.calc)
    moves = [
        dict(type=capi.MOVE_DISPLACEMENT, op=capi.OP_BALL, step=0.1, probability=1.0),
        dict(type=capi.MOVE_DISPLACEMENT, op=capi.OP_BALL, step=0.07, probability=2.0, repeat=2, chain=1),
        dict(type=capi.MOVE_DISPLACEMENT, op=capi.OP_BOX, step=0.03, chain=0),
    ]
    reps, otr = oracle_run(par, pos, np.full(R, n), cell, (1, 1, 1), moves, 99, n_attempts, range(R), temperature=600.0)
    out = mc.download()
    for r in range(R):
        assert np.array_equal(tr["move"][r], otr[r]["move"])
        assert np.array_equal(tr["accepted"][r], otr[r]["accepted"]), f"accept/reject sequence differs for replica {r}"
        assert np.array_equal(tr["moved"][r], otr[r]["moved"])
        np.testing.assert_allclose(tr["energy"][r], otr[r]["energy"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out.positions[r], reps[r].positions, rtol=0, atol=1e-12)
    cnt = mc.counters()
    assert np.array_equal(cnt[:, 0, 0] + cnt[:, 1, 0], np.full(R, n_attempts)) and np.all(cnt[:, 2] == 0)
    assert np.array_equal(cnt[:, 1, 0], np.array([(t["move"] == 1).sum() for t in otr]))
    assert np.array_equal(cnt[:, 1, 1], np.array([((t["move"] == 1) & (t["accepted"] == 1)).sum() for t in otr]))
    # the tracked energy still equals a fresh full evaluation after composite commits and restores
    tracked = mc.batch.energy.cpu().numpy().copy()
    np.testing.assert_allclose(tracked, mc.batch.energy_full().cpu().numpy(), rtol=1e-11, atol=1e-12)


def test_composite_and_label_groups_with_lennard_jones(cuda_lib):
    """The composite and label-group paths on the pair potential (no embedding caches), non-periodic cluster, against the oracle."""
    from quansino_b200 import operations
    from quansino_b200.mc import Canonical
    from quansino_b200.mc.criteria import CanonicalCriteria
    from quansino_b200.moves import DisplacementMove

    R, n_max, n_att = 4, 64, 160
    atoms, pos, cell, n = _gc_setup(R, n_max, 0.0)
    groups = np.concatenate([np.repeat(np.arange(10), 2) * 5 + 3, np.full(n - 20, -1)])  # ten dimers, the rest frozen
    groups[30:36] = 1  # and one group of six with the smallest label
    lab = np.arange(n)
    mc = Canonical(atoms, temperature=300.0, seed=321, trace=True, max_cycles=n_att, resync_interval=0,
                   default_displacement_move=DisplacementMove(groups, operations.Ball(0.08)))  # fmt: skip
    mc.add_move(DisplacementMove(lab, operations.Box(0.05)) * 3, criteria=CanonicalCriteria(), name="triple", probability=1.0)
    mc.run(1)
    tr, out = mc.last_trace, mc.download()
    par = LJParams(sigma=2.5, epsilon=0.1)
    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cell, pbc=(0, 0, 0), n_atoms=n, temperature=300.0, mass=195.084)
        lg = np.full(n_max, -1, np.int32)
        lg[:n] = groups
        la = np.full(n_max, -1, np.int32)
        la[:n] = lab
        specs = [
            capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.08, probability=1.0, labels=lg),
            capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BOX, step=0.05, probability=1.0, repeat=3, chain=0, labels=la),
        ]
        ref = capi.mc_run(par, rep, specs, 321, r, 0, n_att)
        assert np.array_equal(tr["move"][r], ref["move"]) and (ref["move"] == 0).sum() > 20 and (ref["move"] == 1).sum() > 20
        assert np.array_equal(tr["accepted"][r], ref["accepted"]), f"accept/reject sequence differs for replica {r}"
        assert np.array_equal(tr["moved"][r], ref["moved"])
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(out.positions[r, :n], rep.positions[:n], rtol=0, atol=1e-11)


def test_forced_moves_match_oracle(cuda_lib):
    """minimum_count (mc/core.py:331-342) over several launches: every step contains its forced moves at the cycles the oracle
    places them, for every replica."""
    from quansino_b200 import operations
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove

    R, M, steps = 6, 25, 6
    pos, cell, n = cu_replicas(3, R, seed=66)
    lab = np.arange(n)
    mc = Canonical(_emt_atoms(pos, cell, n), temperature=500.0, seed=515, trace=True, max_cycles=M, resync_interval=0, steps_per_launch=2,
                   default_displacement_move=DisplacementMove(lab, operations.Ball(0.1)))  # fmt: skip
    mc.add_move(DisplacementMove(lab, operations.Sphere(0.03)), name="sphere", probability=0.0, minimum_count=4)
    mc.add_move(DisplacementMove(lab, operations.Box(0.2)) * 2, criteria=type(mc.moves["sphere"].criteria)(), name="pair", probability=0.1,
                minimum_count=1)  # fmt: skip
    traces = []
    for step in mc.irun(steps):
        for _ in step:
            pass
        traces.append({k: v.copy() for k, v in mc.last_trace.items()})
    move = np.concatenate([t["move"] for t in traces], axis=1)
    accepted = np.concatenate([t["accepted"] for t in traces], axis=1)
    energy = np.concatenate([t["energy"] for t in traces], axis=1)
    par = emt_params("Cu")
    la = lab.astype(np.int32)
    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cell, temperature=500.0)
        specs = [
            capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.1, probability=1.0, max_cycles=M, labels=la.copy()),
            capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_SPHERE, step=0.03, probability=0.0, minimum_count=4, labels=la.copy()),
            capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BOX, step=0.2, probability=0.1, minimum_count=1, repeat=2, labels=la.copy()),
        ]
        ref = capi.mc_run(par, rep, specs, 515, r, 0, steps * M)
        assert np.array_equal(move[r], ref["move"])
        assert np.array_equal(accepted[r], ref["accepted"])
        np.testing.assert_allclose(energy[r], ref["energy"], rtol=1e-10, atol=1e-12)
        per_step = move[r].reshape(steps, M)
        assert np.all((per_step == 1).sum(axis=1) == 4)  # probability 0: exactly the forced ones
        assert np.all((per_step == 2).sum(axis=1) >= 1)
    assert len({tuple(np.where(move[r].reshape(steps, M)[0] == 1)[0]) for r in range(R)}) > 1  # placements differ between replicas


def test_c_abi_exchange_and_reduction_single_rank(cuda_lib):
    """`qmc_pt_exchange` / `qmc_reduce_observables` without a communicator (one rank): the exchange equals `qmc_pt_swap` on the
    local arrays, the reduction is the identity.  (Two ranks over NCCL: tests/nccl_c_abi.py, run under torchrun.)"""
    import torch

    from quansino_b200 import _lib

    lib = cuda_lib
    n = 16
    rng = np.random.default_rng(8)
    temps = 300.0 * 2.0 ** (rng.permutation(n) / (n - 1))
    energies = 1.0 + 0.01 * temps + rng.normal(scale=0.2, size=n)
    t_a, t_b = torch.from_numpy(temps.copy()).cuda(), torch.from_numpy(temps.copy()).cuda()
    e = torch.from_numpy(energies).cuda()
    scratch = torch.zeros(2 * n, dtype=torch.float64, device="cuda")
    acc_a, acc_b = torch.zeros(n, dtype=torch.int8, device="cuda"), torch.zeros(n, dtype=torch.int8, device="cuda")
    for rnd in range(4):
        _lib.check(lib.qmc_pt_exchange(None, t_a.data_ptr(), e.data_ptr(), n, 99, rnd, scratch.data_ptr(), acc_a.data_ptr(), None))
        _lib.check(lib.qmc_pt_swap(t_b.data_ptr(), e.data_ptr(), n, 99, rnd, acc_b.data_ptr(), None))
        assert torch.equal(t_a, t_b) and torch.equal(acc_a, acc_b)
    assert not np.array_equal(t_a.cpu().numpy(), temps)
    v = torch.arange(5, dtype=torch.float64, device="cuda")
    _lib.check(lib.qmc_reduce_observables(None, v.data_ptr(), 5, None))
    assert torch.equal(v.cpu(), torch.arange(5, dtype=torch.float64))


def test_host_entry_point_with_a_shared_label_array(cuda_lib):
    """qmc_sweep_host with exchange moves: a displacement and an exchange move that share ONE host label array (uploaded once,
    re-labelled once per accepted exchange) give exactly what two equal arrays give, and what the oracle gives."""
    import ctypes as C

    from quansino_b200 import LennardJones, _lib
    from quansino_b200.systems import octahedron, rattle

    R, n_max, n_attempts, mu = 3, 64, 600, -0.5
    pos0, cell = octahedron("Pt", 4, 10.0)
    n = len(pos0)
    pos = np.zeros((R, n_max, 3))
    for r in range(R):
        pos[r, :n] = rattle(pos0, 0.02, 600 + r)
    pot = LennardJones(sigma=2.5, epsilon=0.1).lower("Pt")
    volume = float(np.prod(np.diag(cell)))

    def run(shared: bool):
        hp = np.ascontiguousarray(pos.transpose(0, 2, 1))
        hc = np.ascontiguousarray(np.broadcast_to(cell.reshape(9), (R, 9)))
        n_atoms, n_ex = np.full(R, n, np.int32), np.full(R, n, np.int32)
        energy, temp = np.zeros(R), np.full(R, 300.0)
        counters, status = np.zeros((R, 4, 3), np.int64), np.zeros(R, np.int32)
        lab_a = np.full((R, n_max), -1, np.int32)
        lab_a[:, :n] = np.arange(n)
        lab_b = lab_a if shared else lab_a.copy()
        acc, nat = np.zeros((R, n_attempts), np.int8), None
        b = _lib.Batch()
        b.n_replicas, b.n_max, b.replica_offset = R, n_max, 0
        b.pbc = (C.c_int32 * 3)(0, 0, 0)
        b.pos, b.cell, b.n_atoms, b.energy, b.temperature = (a.ctypes.data for a in (hp, hc, n_atoms, energy, temp))
        b.counters, b.status, b.n_exchange = counters.ctypes.data, status.ctypes.data, n_ex.ctypes.data
        b.mass = b.exchange_mass = 195.084
        b.chemical_potential, b.accessible_volume = mu, volume
        mv = (_lib.Move * 2)()
        mv[0].type, mv[0].op, mv[0].step, mv[0].probability, mv[0].default_label = _lib.MOVE_DISPLACEMENT, _lib.OP_BALL, 0.1, 0.5, _lib.LABEL_NONE
        mv[1].type, mv[1].op, mv[1].probability, mv[1].bias_insert, mv[1].default_label = _lib.MOVE_EXCHANGE, _lib.OP_TRANSLATION, 0.5, 0.5, _lib.LABEL_NONE
        mv[0].labels, mv[1].labels = lab_a.ctypes.data, lab_b.ctypes.data
        tr = _lib.Trace()
        tr.accepted = acc.ctypes.data
        _lib.check(cuda_lib.qmc_sweep_host(C.byref(pot), C.byref(b), mv, 2, 41, 0, n_attempts, C.byref(tr)))
        assert not status.any()
        return hp, n_atoms, n_ex, lab_a, lab_b, acc, energy

    a, s = run(False), run(True)
    for x, y in zip(a, s):
        assert np.array_equal(x, y)
    par = LJParams(sigma=2.5, epsilon=0.1)
    hp, n_atoms, n_ex, lab_a, lab_b, acc, energy = s
    changed = 0
    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cell, pbc=(0, 0, 0), n_atoms=n, temperature=300.0, mass=195.084, chemical_potential=mu,
                           exchange_mass=195.084, n_exchange=n, accessible_volume=volume)  # fmt: skip
        ld = np.full(n_max, -1, np.int32)
        ld[:n] = np.arange(n)
        lx = ld.copy()
        specs = [capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.1, probability=0.5, labels=ld),
                 capi.MoveSpec(capi.MOVE_EXCHANGE, capi.OP_TRANSLATION, probability=0.5, labels=lx)]  # fmt: skip
        ref = capi.mc_run(par, rep, specs, 41, r, 0, n_attempts)
        assert np.array_equal(acc[r], ref["accepted"])
        assert n_atoms[r] == rep.n_atoms and n_ex[r] == rep.n_exchange
        assert np.array_equal(lab_a[r, : rep.n_atoms], ld[: rep.n_atoms]) and np.array_equal(lab_b[r, : rep.n_atoms], lx[: rep.n_atoms])
        assert abs(energy[r] - rep.last_energy) <= 1e-10 * max(1.0, abs(rep.last_energy))
        changed += int((np.diff(np.concatenate([[n], ref["n_atoms"]])) != 0).sum())
    assert changed > 5  # insertions and deletions did happen (and every deletion re-labels the shared table once)
