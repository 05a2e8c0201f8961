"""Parity of the CUDA canonical path against the CPU oracle (config C2 of BASELINE.json at test size).

Bar (north_star): energies within 1e-10 relative, identical accept / reject sequences, bit-exact integer state."""

import numpy as np
import pytest

from oracle import capi
from tests.helpers import cu_replicas, oracle_par, oracle_run

pytestmark = pytest.mark.gpu

RTOL_E = 1e-10


def _driver(pos, cell, n, op, temperature=300.0, seed=1234, trace=True, **kw):
    from quansino_b200 import EMT, BatchedAtoms
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove

    atoms = BatchedAtoms(pos.copy(), cell, np.full(len(pos), n, dtype=np.int32), "Cu", (True, True, True), EMT())
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
