"""Shared helpers of the parity tests: run the CUDA path and the CPU oracle on the same seeded inputs."""

from __future__ import annotations

import numpy as np

from oracle import capi
from oracle.potentials import LJParams, emt_params
from quansino_b200.systems import fcc_cubic, rattle


def cu_replicas(repeat: int, n_replicas: int, stdev: float = 0.05, seed: int = 7, n_max: int | None = None):
    """``n_replicas`` independently rattled fcc Cu cells: positions (R, n_max, 3), cell (3, 3), n_atoms."""
    pos0, cell = fcc_cubic("Cu", repeat)
    n = len(pos0)
    n_max = n if n_max is None else n_max
    pos = np.zeros((n_replicas, n_max, 3))
    for r in range(n_replicas):
        pos[r, :n] = rattle(pos0, stdev, seed * 1000 + r)
    return pos, cell, n


def oracle_par(calc, symbol="Cu"):
    from quansino_b200.calculators import EMT

    if isinstance(calc, EMT):
        return emt_params(symbol)
    return LJParams(sigma=calc.sigma, epsilon=calc.epsilon, rc=calc.rc)


def oracle_run(par, positions, n_atoms, cell, pbc, moves, seed, n_attempts, replica_ids, attempt_base=0, threads=1, **state):
    """Run the C oracle replica by replica (``threads`` > 1: replicas in parallel, the C call releases the GIL).  ``moves`` is a
    list of dicts (MoveSpec kwargs); labels default to arange.  Returns (list of Replica, list of traces)."""
    replica_ids = list(replica_ids)
    if threads > 1 and len(replica_ids) > 1:
        from concurrent.futures import ThreadPoolExecutor

        def one(k):
            sub = {key: (val[k : k + 1] if isinstance(val, np.ndarray) and val.ndim >= 1 and key in ("temperature", "n_exchange") else val) for key, val in state.items()}
            r, t = oracle_run(par, positions[k : k + 1], n_atoms[k : k + 1], cell[k : k + 1] if np.ndim(cell) == 3 else cell, pbc, moves, seed,
                              n_attempts, [replica_ids[k]], attempt_base, **sub)  # fmt: skip
            return r[0], t[0]

        with ThreadPoolExecutor(threads) as ex:
            out = list(ex.map(one, range(len(replica_ids))))
        return [o[0] for o in out], [o[1] for o in out]
    reps, traces = [], []
    for k, rid in enumerate(replica_ids):
        n_max = positions.shape[1]
        specs = []
        for m in moves:
            m = dict(m)
            lab = m.pop("labels", None)
            if m["type"] != capi.MOVE_CELL and m["type"] != capi.MOVE_HMC:
                if lab is None:
                    lab = np.arange(n_max, dtype=np.int32)
                    lab[n_atoms[k] :] = -1
                lab = np.ascontiguousarray(np.asarray(lab, dtype=np.int32)[k] if np.ndim(lab) == 2 else np.asarray(lab, dtype=np.int32)).copy()
                if len(lab) < n_max:
                    lab = np.concatenate([lab, np.full(n_max - len(lab), -1, dtype=np.int32)])
            specs.append(capi.MoveSpec(labels=lab, **m))
        st = {key: (val[k] if isinstance(val, np.ndarray) and val.ndim >= 1 and key in ("temperature", "n_exchange") else val) for key, val in state.items()}
        rep = capi.Replica(positions=positions[k].copy(), cell=cell[k] if np.ndim(cell) == 3 else cell, pbc=pbc, n_atoms=int(n_atoms[k]), **st)
        tr = capi.mc_run(par, rep, specs, seed, int(rid), attempt_base, n_attempts)
        rep.move_specs = specs
        reps.append(rep)
        traces.append(tr)
    return reps, traces
