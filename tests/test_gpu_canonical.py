"""Parity of the CUDA canonical path against the CPU oracle (config C2 of BASELINE.json at test size).

Bar (north_star): energies within 1e-10 relative, identical accept / reject sequences, bit-exact integer state."""

import os

import numpy as np
import pytest

from oracle import capi
from oracle.potentials import emt_params
from tests.helpers import cu_replicas, oracle_par, oracle_run

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_variant")]

RTOL_E = 1e-10


def _driver(pos, cell, n, op, temperature=300.0, seed=1234, trace=True, **kw):
    from quansino_b200 import EMT, BatchedAtoms
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove

    atoms = BatchedAtoms(pos.copy(), cell, np.full(len(pos), n, dtype=np.int32), "Cu", (True, True, True), EMT())
    import os

    kw.setdefault("neighbour_rows", os.environ.get("QMC_FORCE_KERNEL") == "rows")
    mc = Canonical(atoms, temperature=temperature, seed=seed, trace=trace, default_displacement_move=DisplacementMove(np.arange(n), op), **kw)
    return mc


def test_full_energy_matches_oracle(cuda_lib):
    from quansino_b200 import EMT, BatchedAtoms, ReplicaBatch
    import torch

    pos, cell, n = cu_replicas(3, 6)
    atoms = BatchedAtoms(pos, cell, np.full(6, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    batch = ReplicaBatch(atoms, temperature=300.0)
    pairs = torch.zeros(6, dtype=torch.int64, device="cuda")
    e = batch.energy_full(pairs).cpu().numpy()
    sig = batch.sigma1.cpu().numpy()
    par = oracle_par(EMT())
    for r in range(6):
        ref = capi.energy(par, pos[r], cell, (1, 1, 1), want_sigma=True)
        assert abs(e[r] - ref["energy"]) <= RTOL_E * abs(ref["energy"])
        np.testing.assert_allclose(sig[r], ref["sigma1"], rtol=1e-12)
        assert int(pairs[r]) == ref["n_pairs"]


@pytest.mark.parametrize("opname", ["Ball", "Box", "Sphere"])
def test_sweep_matches_oracle(cuda_lib, opname):
    from quansino_b200 import operations

    R, sweeps = 8, 3
    pos, cell, n = cu_replicas(3, R)
    op = getattr(operations, opname)(0.1)
    mc = _driver(pos, cell, n, op, steps_per_launch=sweeps, resync_interval=0)
    mc.run(sweeps)
    tr = mc.last_trace
    par = oracle_par(mc.batch.calc)
    code = {"Ball": capi.OP_BALL, "Box": capi.OP_BOX, "Sphere": capi.OP_SPHERE}[opname]
    reps, otr = oracle_run(
        par, pos, np.full(R, n), cell, (1, 1, 1), [dict(type=capi.MOVE_DISPLACEMENT, op=code, step=0.1)], 1234, sweeps * n, range(R),
        temperature=300.0,
    )  # fmt: skip
    out = mc.download()
    e_gpu = mc.batch.energy.cpu().numpy()
    for r in range(R):
        assert np.array_equal(tr["accepted"][r], otr[r]["accepted"]), f"accept/reject sequence differs for replica {r}"
        assert np.array_equal(tr["moved"][r], otr[r]["moved"])
        np.testing.assert_allclose(tr["energy"][r], otr[r]["energy"], rtol=RTOL_E, atol=1e-12)
        np.testing.assert_allclose(tr["delta"][r], otr[r]["delta"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(out.positions[r], reps[r].positions, rtol=0, atol=1e-12)
        assert abs(e_gpu[r] - reps[r].last_energy) <= RTOL_E * abs(reps[r].last_energy) + 1e-12
    cnt = mc.counters()
    assert np.array_equal(cnt[:, 0, 0], np.full(R, sweeps * n))
    assert np.array_equal(cnt[:, 0, 1], np.array([t["accepted"].sum() for t in otr]))


def test_local_delta_consistent_with_full_recompute(cuda_lib):
    """After many sweeps the incrementally tracked energy equals a fresh full evaluation."""
    from quansino_b200 import operations

    pos, cell, n = cu_replicas(3, 64)
    mc = _driver(pos, cell, n, operations.Ball(0.1), trace=False, steps_per_launch=10, resync_interval=0)
    mc.run(20)
    tracked = mc.batch.energy.cpu().numpy().copy()
    sig = mc.batch.sigma1.cpu().numpy().copy()
    fresh = mc.batch.energy_full().cpu().numpy()
    np.testing.assert_allclose(tracked, fresh, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(sig, mc.batch.sigma1.cpu().numpy(), rtol=1e-12)


def test_replica_offset_independence(cuda_lib):
    """Sharding: replicas 4..7 run alone (replica_offset=4) reproduce replicas 4..7 of the full batch bit for bit."""
    from quansino_b200 import operations

    pos, cell, n = cu_replicas(3, 8)
    full = _driver(pos, cell, n, operations.Ball(0.1), resync_interval=0)
    full.run(2)
    part = _driver(pos[4:], cell, n, operations.Ball(0.1), replica_offset=4, resync_interval=0)
    part.run(2)
    assert np.array_equal(full.download().positions[4:], part.download().positions)
    assert np.array_equal(full.batch.energy.cpu().numpy()[4:], part.batch.energy.cpu().numpy())


def test_host_buffer_entry_point(cuda_lib):
    """qmc_sweep_host: the C-ABI call with HOST buffers (what INTEGRATION.md's stub binds), against the oracle."""
    import ctypes as C

    from quansino_b200 import EMT, _lib

    R = 3
    pos, cell, n = cu_replicas(3, R, seed=13)
    hp = np.ascontiguousarray(pos.transpose(0, 2, 1))
    hc = np.ascontiguousarray(np.broadcast_to(cell.reshape(9), (R, 9)))
    n_atoms = np.full(R, n, np.int32)
    energy = np.zeros(R)
    temp = np.full(R, 300.0)
    counters = np.zeros((R, 4, 3), np.int64)
    status = np.zeros(R, np.int32)
    acc = np.zeros((R, n), np.int8)
    etr = np.zeros((R, n))
    b = _lib.Batch()
    b.n_replicas, b.n_max, b.replica_offset = R, n, 0
    b.pbc = (C.c_int32 * 3)(1, 1, 1)
    b.pos, b.cell, b.n_atoms, b.energy, b.temperature = (a.ctypes.data for a in (hp, hc, n_atoms, energy, temp))
    b.counters, b.status = counters.ctypes.data, status.ctypes.data
    mv = (_lib.Move * 1)()
    mv[0].type, mv[0].op, mv[0].step, mv[0].probability, mv[0].default_label = _lib.MOVE_DISPLACEMENT, _lib.OP_BALL, 0.1, 1.0, _lib.LABEL_NONE
    tr = _lib.Trace()
    tr.accepted, tr.energy = acc.ctypes.data, etr.ctypes.data
    pot = EMT().lower("Cu")
    _lib.check(cuda_lib.qmc_sweep_host(C.byref(pot), C.byref(b), mv, 1, 555, 0, n, C.byref(tr)))
    assert not status.any()
    par = oracle_par(EMT())
    reps, otr = oracle_run(par, pos, n_atoms, cell, (1, 1, 1), [dict(type=capi.MOVE_DISPLACEMENT, op=capi.OP_BALL, step=0.1)], 555, n, range(R), temperature=300.0)
    for r in range(R):
        assert np.array_equal(acc[r], otr[r]["accepted"])
        np.testing.assert_allclose(etr[r], otr[r]["energy"], rtol=RTOL_E, atol=1e-12)
        np.testing.assert_allclose(hp[r].T, reps[r].positions, rtol=0, atol=1e-12)
        assert counters[r, 0, 0] == n and counters[r, 0, 1] == otr[r]["accepted"].sum()


def test_unsupported_requests_are_refused(cuda_lib):
    from quansino_b200 import EMT, BatchedAtoms
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove
    from quansino_b200.systems import fcc_cubic

    pos, cell = fcc_cubic("Cu", 3)
    atoms = BatchedAtoms.replicate(pos, cell, 2, calc=EMT())
    mc = Canonical(atoms, temperature=300.0)
    with pytest.raises(ValueError):  # mc/core.py:161-165: more forced moves than cycles
        mc.add_move(DisplacementMove(np.arange(108)), minimum_count=1000)
    with pytest.raises(ValueError):
        mc.add_move(object())
    with pytest.raises(ZeroDivisionError):
        Canonical(atoms, temperature=0.0)
    pos1, cell1 = fcc_cubic("Cu", 1)  # the reference's 4-atom fixture: cell shorter than the EMT cutoff
    with pytest.raises(NotImplementedError):
        Canonical(BatchedAtoms.replicate(pos1, cell1, 2, calc=EMT()), temperature=300.0)


def test_ragged_batch_and_frozen_atoms(cuda_lib):
    """Replicas with different atom counts (padded rows) and atoms excluded by negative labels, against the oracle."""
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove

    R, n_max = 5, 108
    pos, cell, n = cu_replicas(3, R, seed=23)
    n_atoms = np.array([108, 107, 100, 108, 64], dtype=np.int32)  # vacancies: trailing atoms removed
    labels = np.tile(np.arange(n_max), (R, 1))
    labels[:, ::7] = -1  # every 7th atom is frozen
    for r in range(R):
        pos[r, n_atoms[r] :] = 0.0
    atoms = BatchedAtoms(pos.copy(), cell, n_atoms, "Cu", (True, True, True), EMT())
    mc = Canonical(atoms, temperature=600.0, seed=31, trace=True, max_cycles=150, resync_interval=0,
                   default_displacement_move=DisplacementMove(labels, operations.Ball(0.15)))  # fmt: skip
    mc.run(1)
    tr = mc.last_trace
    out = mc.download()
    par = oracle_par(EMT())
    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cell, n_atoms=int(n_atoms[r]), temperature=600.0)
        lab = labels[r].astype(np.int32).copy()
        lab[n_atoms[r] :] = -1
        ref = capi.mc_run(par, rep, [capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.15, labels=lab)], 31, r, 0, 150)
        assert np.array_equal(tr["accepted"][r], ref["accepted"])
        assert np.array_equal(tr["moved"][r], ref["moved"])
        assert not np.any(ref["moved"] % 7 == 0)  # frozen atoms never move
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=RTOL_E, atol=1e-12)
        np.testing.assert_allclose(out.positions[r, : n_atoms[r]], rep.positions[: n_atoms[r]], rtol=0, atol=1e-12)
    assert np.array_equal(out.n_atoms, n_atoms)


def test_label_groups_move_rigidly(cuda_lib):
    """Per-replica label tables with groups of different sizes, ragged atom counts and frozen atoms, against the oracle."""
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove

    R, n_max, n_att = 4, 108, 120
    pos, cell, n = cu_replicas(3, R, seed=29)
    n_atoms = np.array([108, 108, 96, 70], dtype=np.int32)
    rng = np.random.default_rng(5)
    labels = np.full((R, n_max), -1)
    for r in range(R):
        labels[r, : n_atoms[r]] = rng.integers(-1, 25 + 10 * r, size=n_atoms[r]) * 3  # sparse, unordered, repeated label values
        pos[r, n_atoms[r] :] = 0.0
    atoms = BatchedAtoms(pos.copy(), cell, n_atoms, "Cu", (True, True, True), EMT())
    mc = Canonical(atoms, temperature=500.0, seed=47, trace=True, max_cycles=n_att, resync_interval=0,
                   default_displacement_move=DisplacementMove(labels, operations.Box(0.05)))  # fmt: skip
    mc.run(1)
    tr, out, par = mc.last_trace, mc.download(), oracle_par(EMT())
    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cell, n_atoms=int(n_atoms[r]), temperature=500.0)
        lab = np.where(labels[r] < 0, -1, labels[r]).astype(np.int32)
        ref = capi.mc_run(par, rep, [capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BOX, step=0.05, labels=lab)], 47, r, 0, n_att)
        assert np.array_equal(tr["accepted"][r], ref["accepted"])
        assert np.array_equal(tr["moved"][r], ref["moved"])  # first member of the displaced group
        np.testing.assert_allclose(tr["energy"][r], ref["energy"], rtol=RTOL_E, atol=1e-12)
        np.testing.assert_allclose(out.positions[r, : n_atoms[r]], rep.positions[: n_atoms[r]], rtol=0, atol=1e-12)
        moved_rows = np.any(np.abs(out.positions[r, : n_atoms[r]] - pos[r, : n_atoms[r]]) > 0, axis=1)
        assert not np.any(moved_rows & (lab[: n_atoms[r]] < 0))  # frozen atoms never move
    tracked = mc.batch.energy.cpu().numpy().copy()
    np.testing.assert_allclose(tracked, mc.batch.energy_full().cpu().numpy(), rtol=1e-11, atol=1e-12)


def test_no_movable_atoms_means_failed_moves(cuda_lib):
    """All labels negative: every attempt fails (the reference's None), nothing changes, failed counter counts."""
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove

    pos, cell, n = cu_replicas(3, 2, seed=29)
    atoms = BatchedAtoms(pos.copy(), cell, np.full(2, n, np.int32), "Cu", (True, True, True), EMT())
    mc = Canonical(atoms, temperature=300.0, seed=1, trace=True, max_cycles=20, resync_interval=0,
                   default_displacement_move=DisplacementMove(np.full(n, -1), operations.Ball(0.1)))  # fmt: skip
    e0 = mc.batch.energy_full().cpu().numpy().copy()
    mc.run(1)
    assert np.all(mc.last_trace["accepted"] == -1)
    assert np.array_equal(mc.download().positions, pos)
    assert np.array_equal(mc.counters()[:, 0], np.array([[20, 0, 20], [20, 0, 20]]))
    assert np.array_equal(mc.batch.energy.cpu().numpy(), e0)
    assert np.allclose(mc.update_acceptance_rate(), 0.0)


def test_lennard_jones_periodic_canonical(cuda_lib):
    """The pair potential on a periodic cell whose edge is below twice the cutoff (second images needed)."""
    from quansino_b200 import BatchedAtoms, LennardJones, operations
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove
    from oracle.potentials import LJParams

    R = 4
    pos, cell, n = cu_replicas(3, R, seed=37)
    calc = LennardJones(sigma=2.3, epsilon=0.2, rc=5.6)  # L = 10.83 < 2 * 5.6
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n, np.int32), "Cu", (True, True, True), calc)
    mc = Canonical(atoms, temperature=500.0, seed=41, trace=True, max_cycles=200, resync_interval=0,
                   default_displacement_move=DisplacementMove(np.arange(n), operations.Box(0.12)))  # fmt: skip
    mc.run(1)
    tr = mc.last_trace
    par = LJParams(sigma=2.3, epsilon=0.2, rc=5.6)
    reps, otr = oracle_run(par, pos, np.full(R, n), cell, (1, 1, 1), [dict(type=capi.MOVE_DISPLACEMENT, op=capi.OP_BOX, step=0.12)], 41, 200,
                           range(R), temperature=500.0, mass=63.546)  # fmt: skip
    for r in range(R):
        assert np.array_equal(tr["accepted"][r], otr[r]["accepted"])
        np.testing.assert_allclose(tr["energy"][r], otr[r]["energy"], rtol=RTOL_E, atol=1e-11)


def test_overflow_is_reported_like_the_reference(cuda_lib):
    """A hugely favourable move makes exp(-dE/kT) overflow: the reference raises OverflowError (mc/criteria.py:108)."""
    from quansino_b200 import BatchedAtoms, LennardJones, operations
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove

    pos = np.zeros((1, 2, 3))
    pos[0, 1] = [0.9, 0.0, 0.0]  # two LJ atoms far up the repulsive wall: any move apart releases > 1e4 eV
    cell = np.eye(3) * 40.0
    atoms = BatchedAtoms(pos, cell, np.array([2], np.int32), "Cu", (False, False, False), LennardJones(sigma=2.5, epsilon=1.0))
    mc = Canonical(atoms, temperature=1.0, seed=3, max_cycles=64, resync_interval=0,
                   default_displacement_move=DisplacementMove(np.arange(2), operations.Ball(0.3)))  # fmt: skip
    with pytest.raises(OverflowError):
        mc.run(1)


def test_library_owned_batch_handle(cuda_lib):
    """The handle API (qmc_batch_create / upload / view / download): a host without a device allocator of its own drives the
    same kernels and gets the same bits as the torch-backed path."""
    import ctypes as C

    from quansino_b200 import EMT, _lib, operations

    lib = cuda_lib
    R, sweeps = 5, 2
    pos, cell, n = cu_replicas(3, R, seed=71)
    mc = _driver(pos, cell, n, operations.Ball(0.1), steps_per_launch=sweeps, resync_interval=0)
    mc.run(sweeps)
    want_pos, want_e, want_cnt = mc.download().positions, mc.batch.energy.cpu().numpy(), mc.counters()

    h = C.c_void_p()
    _lib.check(lib.qmc_batch_create(R, n, 0, 0, C.byref(h)))
    try:
        def up(field, arr):
            arr = np.ascontiguousarray(arr)
            _lib.check(lib.qmc_batch_upload(h, field, arr.ctypes.data, arr.nbytes))

        def down(field, arr):
            _lib.check(lib.qmc_batch_download(h, field, arr.ctypes.data, arr.nbytes))
            return arr

        up(_lib.FIELD_POS, pos.transpose(0, 2, 1))
        up(_lib.FIELD_CELL, np.broadcast_to(cell.reshape(9), (R, 9)))
        up(_lib.FIELD_N_ATOMS, np.full(R, n, dtype=np.int32))
        up(_lib.FIELD_TEMPERATURE, np.full(R, 300.0))
        b = _lib.Batch()
        _lib.check(lib.qmc_batch_view(h, C.byref(b)))
        b.pbc = (C.c_int32 * 3)(1, 1, 1)
        b.mass = 63.546
        pot = EMT().lower("Cu")
        e0 = down(_lib.FIELD_ENERGY, np.zeros(R))
        assert np.all(np.isnan(e0))  # not primed yet
        _lib.check(lib.qmc_energy_full(C.byref(pot), C.byref(b), None, None))
        move = mc._lower(["default_displacement_move"])[0]
        _lib.check(lib.qmc_sweep(C.byref(pot), C.byref(b), move, 1, 1234, 0, sweeps * n, None, None))
        got_pos = down(_lib.FIELD_POS, np.zeros((R, 3, n))).transpose(0, 2, 1)
        assert np.array_equal(got_pos, want_pos)
        if os.environ.get("QMC_FORCE_KERNEL") == "rows":  # the torch-backed run read neighbour rows: image vectors round differently (<= 1 ulp)
            np.testing.assert_allclose(down(_lib.FIELD_ENERGY, np.zeros(R)), want_e, rtol=1e-13)
        else:
            assert np.array_equal(down(_lib.FIELD_ENERGY, np.zeros(R)), want_e)
        assert np.array_equal(down(_lib.FIELD_COUNTERS, np.zeros((R, 4, 3), dtype=np.int64)), want_cnt)
        assert lib.qmc_batch_upload(h, _lib.FIELD_POS, pos.ctypes.data, 8) == _lib.ERR_INVALID  # wrong size
        assert lib.qmc_batch_upload(h, _lib.FIELD_MOMENTA, pos.ctypes.data, pos.nbytes) == _lib.ERR_INVALID  # created without momenta
        assert lib.qmc_batch_labels(h, 0) is None
    finally:
        _lib.check(lib.qmc_batch_destroy(h))


def test_seed_zero_is_a_seed_and_tables_are_relowered(cuda_lib):
    """(1) `seed=0` keys the stream like any other seed (the reference: `Generator(PCG64(0))`, mc/driver.py:77-78) instead of being
    replaced by a random one.  (2) Changes made to a move BETWEEN runs -- step size, operation -- reach the kernel: the reference
    re-reads them on every call (moves/displacement.py:109-141)."""
    from quansino_b200 import operations

    R = 3
    pos, cell, n = cu_replicas(3, R, seed=12)
    runs = []
    for _ in range(2):
        mc = _driver(pos, cell, n, operations.Ball(0.1), seed=0, trace=True, resync_interval=0)
        assert mc._seed == 0
        mc.run(1)
        mc.moves["default_displacement_move"].move.operation.step_size = 0.04
        mc.run(1)
        mc.moves["default_displacement_move"].move.operation = operations.Box(0.05)
        mc.run(1)
        runs.append((mc.download().positions.copy(), mc.last_trace["accepted"].copy()))
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
    par = emt_params("Cu")
    for r in range(R):
        rep = capi.Replica(positions=pos[r].copy(), cell=cell, temperature=300.0)
        lab = np.arange(n, dtype=np.int32)
        for k, (op, step) in enumerate(((capi.OP_BALL, 0.1), (capi.OP_BALL, 0.04), (capi.OP_BOX, 0.05))):
            tr = capi.mc_run(par, rep, [capi.MoveSpec(capi.MOVE_DISPLACEMENT, op, step=step, labels=lab)], 0, r, k * n, n)
        assert np.array_equal(runs[0][1][r], tr["accepted"])
        np.testing.assert_allclose(runs[0][0][r], rep.positions, rtol=0, atol=1e-12)


def test_device_emt_reproduces_ases_atomization_tutorial(cuda_lib):
    """ASE's documentation prints 5.10 / 0.44 / 9.76 eV for the N atom, the N2 molecule (d = 1.1 A) and the atomization energy with
    `EMT()`: the device kernels (sweep-path energy kernel and the Verlet-list force kernels) give the same digits."""
    from quansino_b200 import EMT, BatchedAtoms, ReplicaBatch

    pos = np.zeros((2, 2, 3))
    pos[0, 1, 2] = 1.1  # replica 0: the molecule; replica 1: one atom
    atoms = BatchedAtoms(pos, np.eye(3) * 10.0, np.array([2, 1], dtype=np.int32), "N", (False, False, False), EMT())
    batch = ReplicaBatch(atoms, EMT())
    e = batch.energy_full().cpu().numpy()
    assert f"{e[1]:.2f}" == "5.10" and f"{e[0]:.2f}" == "0.44" and f"{2 * e[1] - e[0]:.2f}" == "9.76"
    assert abs(e[0] - 0.440343573035614) <= 1e-12
    f = batch.forces_full().cpu().numpy()
    assert abs(batch.energy.cpu().numpy()[0] - e[0]) <= 1e-12 and abs(f[0, 2, 0] + f[0, 2, 1]) <= 1e-12 and abs(f[0, 2, 1] + 3.25179997) < 1e-6
