"""The C oracle against the golden traces of the GENUINE quansino sources (tests/golden/, oracle/gen_golden.py).

This is what pins the oracle's restatement of quansino's own semantics: move selection, proposals, criteria,
save / revert, label / index bookkeeping, Verlet.  (The ASE arithmetic underneath stays "parity unpinned": the golden
runs use the numpy restatement of ASE's EMT / LJ, see oracle/__init__.py.)"""

import numpy as np
import pytest

from tests.golden_cases import CASES, load, oracle_replay


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_genuine_quansino(name):
    g = load(name)
    rep, tr, labels = oracle_replay(g)
    assert np.array_equal(tr["accepted"], g["accepted"]), "accept / reject / failed sequence differs from quansino"
    assert np.array_equal(tr["move"], g["move"])
    np.testing.assert_allclose(tr["energy"], g["energy"], rtol=1e-10, atol=1e-11)
    assert np.array_equal(tr["n_atoms"], g["n_atoms"])
    n = rep.n_atoms
    np.testing.assert_allclose(rep.positions[:n], g["positions"], rtol=0, atol=1e-11)
    if g["kind"] in ("isobaric", "hybrid", "forced"):
        np.testing.assert_allclose(rep.cell, g["cell"], rtol=1e-14)
    if g["kind"] in ("gc_lj", "gc_emt"):
        assert rep.n_exchange == g["n_exchange"]
        assert np.array_equal(labels["displacement"][:n], g["labels_displacement"])
        assert np.array_equal(labels["exchange"][:n], g["labels_exchange"])
    if g["kind"] == "hmc":
        np.testing.assert_allclose(rep.momenta, g["momenta"], rtol=0, atol=1e-10)
        assert abs(rep.last_kinetic - g["last_kinetic"]) <= 1e-10 * abs(g["last_kinetic"])


def test_golden_cases_exercise_every_branch():
    g = load("gc_lj")
    d = np.diff(np.concatenate([[len(g["positions0"])], g["n_atoms"]]))
    assert (d == 1).sum() > 5 and (d == -1).sum() > 5 and (g["accepted"] == -1).sum() > 0
    g = load("isobaric_pcell")
    assert ((g["move"] == 1) & (g["accepted"] == 1)).sum() > 0 and ((g["move"] == 1) & (g["accepted"] == 0)).sum() > 0


def test_oracle_reproduces_genuine_force_bias():
    """ForceBias.step (mc/fbmc.py:202-251) run by the genuine sources: same positions bit for bit, same energies."""
    from oracle import capi
    from oracle.potentials import emt_params

    g = load("force_bias")
    pos = g["positions0"].copy()
    par = emt_params("Cu")
    for s in range(len(g["energy"])):
        e, gmax, calls = capi.fbmc_step(par, pos, g["cell"], (1, 1, 1), 63.546, g["temperature"], g["delta"], int(g["seed"]), int(g["replica"]), s)
        assert np.array_equal(pos, g["trajectory"][s])
        assert abs(e - g["energy"][s]) <= 1e-11 * abs(g["energy"][s]) and abs(gmax - g["gamma_max"][s]) <= 1e-12 and calls >= 5
