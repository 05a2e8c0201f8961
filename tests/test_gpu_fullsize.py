"""BASELINE.json's headline configuration at FULL size (4096 replicas x 108-atom Cu EMT per GPU), through properties that do not
need the oracle to run the whole batch: sampled replicas against the CPU oracle, tracked energy = fresh full evaluation,
sharding invariance (bit-exact), determinism (bit-exact), counter bookkeeping."""

import numpy as np
import pytest

from oracle import capi
from tests.helpers import cu_replicas, oracle_par, oracle_run

pytestmark = pytest.mark.gpu

R_FULL, SWEEPS, SEED = 4096, 3, 20241


def _mc(pos, cell, n, replica_offset=0, trace=False):
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove

    atoms = BatchedAtoms(pos.copy(), cell, np.full(len(pos), n, dtype=np.int32), "Cu", (True, True, True), EMT())
    return Canonical(atoms, temperature=300.0, seed=SEED, trace=trace, replica_offset=replica_offset, resync_interval=0,
                     steps_per_launch=SWEEPS, default_displacement_move=DisplacementMove(np.arange(n), operations.Ball(0.1)))  # fmt: skip


@pytest.fixture(scope="module")
def full_run(cuda_lib):
    pos, cell, n = cu_replicas(3, R_FULL)
    mc = _mc(pos, cell, n, trace=True)
    mc.run(SWEEPS)
    mc.tracked_energy = mc.batch.energy.cpu().numpy().copy()  # before any test re-primes the caches
    return pos, cell, n, mc


def test_sampled_replicas_match_the_oracle(full_run):
    pos, cell, n, mc = full_run
    ids = [0, 1, 27, 28, 2047, 3000, 4094, 4095]  # first / last warp of a CTA, first / last CTA
    par = oracle_par(mc.batch.calc)
    reps, otr = oracle_run(par, pos[ids], np.full(len(ids), n), cell, (1, 1, 1),
                           [dict(type=capi.MOVE_DISPLACEMENT, op=capi.OP_BALL, step=0.1)], SEED, SWEEPS * n, ids, temperature=300.0)  # fmt: skip
    out = mc.download()
    tr = mc.last_trace
    for k, r in enumerate(ids):
        assert np.array_equal(tr["accepted"][r], otr[k]["accepted"]), f"accept/reject sequence differs for replica {r}"
        assert np.array_equal(tr["moved"][r], otr[k]["moved"])
        np.testing.assert_allclose(tr["energy"][r], otr[k]["energy"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out.positions[r], reps[k].positions, rtol=0, atol=1e-12)


def test_tracked_energy_equals_fresh_evaluation_everywhere(full_run):
    _, _, n, mc = full_run
    tracked = mc.tracked_energy
    sig = mc.batch.sigma1.cpu().numpy().copy()
    fresh = mc.batch.energy_full().cpu().numpy()
    np.testing.assert_allclose(tracked, fresh, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(sig, mc.batch.sigma1.cpu().numpy(), rtol=1e-12)
    cnt = mc.counters()
    assert np.array_equal(cnt[:, 0, 0], np.full(R_FULL, SWEEPS * n)) and np.all(cnt[:, 0, 2] == 0) and np.all(cnt[:, 1:] == 0)
    acc = cnt[:, 0, 1] / (SWEEPS * n)
    assert 0.5 < acc.mean() < 0.85 and acc.min() > 0.3
    assert int(mc.batch.status.max().item()) == 0
    assert np.array_equal(mc.last_trace["accepted"].sum(axis=1), cnt[:, 0, 1])


def test_sharding_and_determinism_are_bit_exact(full_run):
    """Two half batches with their global replica ids reproduce the full batch bit for bit, and so does a rerun."""
    pos, cell, n, mc = full_run
    ref = mc.download().positions
    half = R_FULL // 2
    for lo in (0, half):
        part = _mc(pos[lo : lo + half], cell, n, replica_offset=lo)
        part.run(SWEEPS)
        assert np.array_equal(part.download().positions, ref[lo : lo + half])
    again = _mc(pos, cell, n)
    again.run(SWEEPS)
    assert np.array_equal(again.download().positions, ref)
    assert np.array_equal(again.batch.energy.cpu().numpy(), mc.tracked_energy)


def test_long_run_drift_of_the_tracked_energy(cuda_lib):
    """64 replicas x 300 sweeps (32 400 attempts each) WITHOUT re-priming: the incrementally tracked energy and sigma_1 stay
    within 1e-10 relative of a fresh full evaluation (the driver's default resync_interval = 100 keeps it far below that)."""
    pos, cell, n = cu_replicas(3, 64, seed=77)
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove

    atoms = BatchedAtoms(pos.copy(), cell, np.full(64, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    mc = Canonical(atoms, temperature=600.0, seed=5, resync_interval=0, steps_per_launch=50,
                   default_displacement_move=DisplacementMove(np.arange(n), operations.Ball(0.1)))  # fmt: skip
    mc.run(300)
    tracked = mc.batch.energy.cpu().numpy().copy()
    sig = mc.batch.sigma1.cpu().numpy().copy()
    fresh = mc.batch.energy_full().cpu().numpy()
    drift = np.max(np.abs(tracked - fresh) / np.abs(fresh))
    assert drift < 1e-10, drift
    np.testing.assert_allclose(sig, mc.batch.sigma1.cpu().numpy(), rtol=1e-11)
    assert mc.counters()[:, 0, 0].min() == 300 * n
