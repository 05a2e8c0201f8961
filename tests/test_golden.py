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
    if g["kind"] == "general_cell":  # expm by Jacobi rotations here, by Pade approximants in scipy: equal to rounding
        np.testing.assert_allclose(rep.cell, g["cell"], rtol=0, atol=1e-12)
    if g["kind"] in ("gc_lj", "gc_emt", "gc_molecule", "gc_composite"):
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


@pytest.mark.parametrize("name", ["force_bias", "adaptive_force_bias", "adaptive_force_bias_exp"])
def test_oracle_reproduces_genuine_force_bias(name):
    """ForceBias.step (mc/fbmc.py:202-251) run by the genuine sources: same positions bit for bit, same energies.  The adaptive
    cases (mc/fbmc.py:285-488) replay the deltas the genuine ``update_delta`` chose and check them against its formula."""
    from oracle import capi
    from oracle.potentials import emt_energy_forces, emt_params

    g = load(name)
    if name == "adaptive_force_bias":  # forces scheme, tanh: no committee -> reference variance every step
        assert np.all(g["delta_steps"] == 0.04 + (0.13 - 0.04) * (1 - np.tanh(0.1 / 0.1 * np.arctanh(0.5))))
    if name == "adaptive_force_bias_exp":  # energy scheme, exp: first step falls back, then std(E_atom) / N of the last evaluation
        assert g["delta_steps"][0] == 0.02 + (0.2 - 0.02) * np.exp(-0.3 / 0.3 * np.log(2))
        for s in (1, 2):
            e_atoms = emt_energy_forces(g["trajectory"][s - 1], g["cell"], (True,) * 3, emt_params("Cu"), want_forces=False)["energies"]
            d = 0.02 + (0.2 - 0.02) * np.exp(-(np.std(e_atoms) / len(e_atoms)) / 0.3 * np.log(2))
            assert abs(d - g["delta_steps"][s]) <= 1e-15
    pos = g["positions0"].copy()
    par = emt_params("Cu")
    for s in range(len(g["energy"])):
        e, gmax, calls = capi.fbmc_step(par, pos, g["cell"], (1, 1, 1), 63.546, g["temperature"], float(g["delta_steps"][s]), int(g["seed"]), int(g["replica"]), s)
        assert np.array_equal(pos, g["trajectory"][s])
        assert abs(e - g["energy"][s]) <= 1e-11 * abs(g["energy"][s]) and abs(gmax - g["gamma_max"][s]) <= 1e-12 and calls >= 5


def test_molecular_exchange_fixture_bookkeeping():
    """``gc_molecule``: dimers inserted / deleted by the genuine sources (ExchangeMove + TranslationRotation).  The device path does
    not lower molecular exchange yet (DESIGN.md section 8 has the plan); this pins what the plan relies on: a molecule carries ONE
    label in every move, counts change by whole molecules, and labels stay non-decreasing with the atom index, so that a label
    group is a contiguous run even after deletions have left gaps in the label values."""
    g = load("gc_molecule")
    n_atoms = np.concatenate([[len(g["positions0"])], g["n_atoms"]])
    steps = np.diff(n_atoms)
    assert set(np.unique(steps)) <= {-2, 0, 2} and (steps > 0).sum() > 20 and (steps < 0).sum() > 20
    moved = g["accepted"] == 1
    assert np.all(steps[~moved] == 0) and np.all(steps[(g["move"] == 0)] == 0)
    n0, n = len(g["positions0"]), int(n_atoms[-1])
    assert n == len(g["positions"]) and g["n_exchange"] == (n - n0) // 2
    for key, first in (("labels_exchange", -1), ("labels_displacement", None)):
        lab = g[key]
        assert len(lab) == n
        if first is not None:
            assert np.all(lab[:n0] == first)
        tail = lab[n0:]
        assert np.all(tail[0::2] == tail[1::2]) and np.all(np.diff(tail[0::2]) > 0)  # pairs share a label; labels increase
        assert np.any(np.diff(tail[0::2]) > 1)  # deletions left gaps: the labels are no longer dense
    d = np.linalg.norm(g["positions"][n0::2] - g["positions"][n0 + 1 :: 2], axis=1)
    np.testing.assert_allclose(d, 2.8, rtol=0, atol=1e-12)  # rigid bodies: placed by a rotation, moved as a label group
