"""The CPU oracle against every known answer available for this path."""

import math

import numpy as np
import pytest

from oracle import capi, philox, units
from oracle.potentials import LJParams, emt_energy_forces, emt_params, lj_energy_forces
from quansino_b200.systems import fcc_cubic, octahedron, rattle


# --- Philox4x32-10: Random123 known-answer vectors -------------------------------------------------------------
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


@pytest.mark.parametrize("ctr,key,out", KAT)
def test_philox_known_answers(ctr, key, out):
    assert philox.philox4x32_10(ctr, key) == out
    assert capi.philox(ctr, key) == out


def test_uniform_slots_agree_between_python_and_c():
    for seed, attempt, replica, block in [(1, 0, 0, 0), (2**40 + 17, 2**33 + 5, 4095, 2), (99, 123456, 7, 16)]:
        assert philox.uniform_pair(seed, attempt, replica, block) == capi.uniform_pair(seed, attempt, replica, block)
    u = np.array([capi.uniform_pair(5, a, 1, 1) for a in range(2000)])
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02
    z = np.array([capi.lib().qo_normal(7, 3, 2, m) for m in range(20000)])
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1.0) < 0.03
    assert capi.lib().qo_normal(7, 3, 2, 11) == philox.normal(7, 3, 2, 11)


# --- units and criteria constants pinned by the reference's own tests ------------------------------------------
def test_units_match_ase_codata_2014():
    assert units.kB == 8.617330337217213e-05
    assert units.bar == 6.241509125883258e-07
    assert abs(units.fs - 0.09822694788464065) < 2e-17
    assert units.Bohr == 0.5291772105638411  # derived like ase.units (CODATA value 0.52917721067 to 2e-10)
    assert abs(units.Bohr - 0.52917721067) < 2e-10


def test_grand_canonical_constants_of_the_reference_tests():
    """tests/mc/test_grand_canonical.py:557-600 of the reference: Lambda(H2, 298 K), prefactor and probability."""
    mass_h2 = 2 * 1.008
    lam = capi.lib().qo_debroglie_wavelength(mass_h2, 298.0)
    assert np.allclose(lam, 0.71228)
    # the reference test uses a 4-atom Cu cell (a = 3.61 ... cubic 1x1x1 -> V = 3.61^3) and N_ex = 0 at that point;
    # its numbers: accessible_volume / (Lambda^3 (N + 1)) = 13.83662789
    volume = 13.83662789 * lam**3
    p = capi.lib().qo_gc_acceptance(0.0, 1, 0, volume, mass_h2, 298.0, 0.0)
    assert np.allclose(p, 13.83662789)
    p = capi.lib().qo_gc_acceptance(-9.9, 1, 0, volume, mass_h2, 298.0, -10.0)
    expected = 13.83662789 * math.exp(-10 / (298 * units.kB)) * math.exp(9.9 / (298 * units.kB))
    assert np.allclose(p, expected) and np.allclose(p, 0.28172731, rtol=1e-6)
    # deletion: factorial term N_ex, volume^-1, Lambda^3
    p = capi.lib().qo_gc_acceptance(0.3, -1, 5, 100.0, 63.546, 500.0, -0.2)
    lam = capi.lib().qo_debroglie_wavelength(63.546, 500.0)
    assert np.allclose(p, (1 / 100.0) * 5 * lam**3 * math.exp((0.2 - 0.3) / (500 * units.kB)))


# --- EMT / LJ: known answer and self-consistency (parity unpinned at the ASE boundary, see oracle/__init__.py) ---
def test_emt_bulk_copper_known_answer():
    """ASE's widely quoted bulk('Cu') + EMT() energy: -0.00568151 eV / atom at a = 3.61 A (external anchor)."""
    pos, cell = fcc_cubic("Cu", 3)
    par = emt_params("Cu")
    r = capi.energy(par, pos, cell, (1, 1, 1), want_forces=True, want_sigma=True)
    assert abs(r["energy"] / 108 - (-0.00568151)) < 5e-9
    assert r["n_pairs"] == 108 * 78  # shells 12 + 6 + 24 + 12 + 24; minimum image alone would find 66
    assert np.abs(r["forces"]).max() < 1e-12
    assert abs(par.gamma1 - 1.0241586143594834) < 1e-14 and abs(par.gamma2 - 1.0335943498160205) < 1e-14
    # the primitive 4-atom cell (self-images, many images per pair) gives the same energy per atom
    pos1, cell1 = fcc_cubic("Cu", 1)
    assert abs(capi.energy(par, pos1, cell1, (1, 1, 1))["energy"] / 4 - r["energy"] / 108) < 1e-13
    # an isolated atom: -E0
    assert capi.energy(par, np.zeros((1, 3)), np.eye(3) * 30, (0, 0, 0))["energy"] == 3.51


def test_c_and_numpy_restatements_agree():
    par = emt_params("Cu")
    pos, cell = fcc_cubic("Cu", 3)
    pos = rattle(pos, 0.08, 5)
    pos[7] += [11.0, -21.7, 0.3]  # an atom far outside the cell: images must absorb it
    a = capi.energy(par, pos, cell, (1, 1, 1), want_forces=True, want_sigma=True)
    b = emt_energy_forces(pos, cell)
    assert abs(a["energy"] - b["energy"]) < 1e-12
    assert a["n_pairs"] == b["n_pairs"]
    np.testing.assert_allclose(a["forces"], b["forces"], atol=1e-12)
    np.testing.assert_allclose(a["sigma1"], b["sigma1"], rtol=1e-13)
    lj = LJParams(sigma=2.5, epsilon=0.1)
    p, c = octahedron("Pt", 4, 10.0)
    a = capi.energy(lj, p, c, (0, 0, 0), want_forces=True)
    b = lj_energy_forces(p, c, (0, 0, 0), lj)
    assert abs(a["energy"] - b["energy"]) < 1e-12 and a["n_pairs"] == b["n_pairs"]
    np.testing.assert_allclose(a["forces"], b["forces"], atol=1e-12)
    lj2 = LJParams(sigma=2.3, epsilon=0.2)
    pos2, cell2 = fcc_cubic("Cu", 2)
    a = capi.energy(lj2, rattle(pos2, 0.05, 2), cell2, (1, 1, 1))
    b = lj_energy_forces(rattle(pos2, 0.05, 2), cell2, (1, 1, 1), lj2)
    assert abs(a["energy"] - b["energy"]) < 1e-11 and a["n_pairs"] == b["n_pairs"]


@pytest.mark.parametrize("kind", ["emt", "lj"])
def test_forces_are_the_gradient(kind):
    pos, cell = fcc_cubic("Cu", 2)
    pos = rattle(pos, 0.05, 9)
    par = emt_params("Cu") if kind == "emt" else LJParams(sigma=2.3, epsilon=0.2)
    f = capi.energy(par, pos, cell, (1, 1, 1), want_forces=True)["forces"]
    h = 1e-5
    for i, x in [(0, 0), (5, 1), (17, 2)]:
        p = pos.copy()
        p[i, x] += h
        ep = capi.energy(par, p, cell, (1, 1, 1))["energy"]
        p[i, x] -= 2 * h
        em = capi.energy(par, p, cell, (1, 1, 1))["energy"]
        assert abs(-(ep - em) / (2 * h) - f[i, x]) < 1e-7


def test_energy_invariances():
    par = emt_params("Cu")
    pos, cell = fcc_cubic("Cu", 3)
    pos = rattle(pos, 0.05, 3)
    e = capi.energy(par, pos, cell, (1, 1, 1))["energy"]
    assert abs(capi.energy(par, pos + [0.37, -1.2, 5.0], cell, (1, 1, 1))["energy"] - e) < 1e-11
    perm = np.random.default_rng(0).permutation(len(pos))
    assert abs(capi.energy(par, pos[perm], cell, (1, 1, 1))["energy"] - e) < 1e-11


# --- Monte Carlo semantics -------------------------------------------------------------------------------------
def test_rejected_moves_restore_bit_exact_state_and_energy_is_consistent():
    par = emt_params("Cu")
    pos, cell = fcc_cubic("Cu", 3)
    pos = rattle(pos, 0.05, 1)
    rep = capi.Replica(positions=pos.copy(), cell=cell, temperature=300.0)
    spec = [capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.1, labels=np.arange(108, dtype=np.int32))]
    before = rep.positions.copy()
    tr = capi.mc_run(par, rep, spec, 42, 0, 0, 60)
    moved_accepted = set(tr["moved"][tr["accepted"] == 1].tolist())
    untouched = [i for i in range(108) if i not in moved_accepted]
    assert np.array_equal(rep.positions[untouched], before[untouched])  # bit-exact
    assert rep.last_energy == capi.energy(par, rep.positions, cell, (1, 1, 1))["energy"]
    # P and M of the local algorithm: ~156 pair and ~67 embedding evaluations per attempt
    assert 140 < rep.counters[1] / 60 < 170 and 55 < rep.counters[2] / 60 < 80


def test_negative_labels_and_empty_label_set():
    par = emt_params("Cu")
    pos, cell = fcc_cubic("Cu", 2)
    lab = np.full(32, -1, dtype=np.int32)
    rep = capi.Replica(positions=pos.copy(), cell=cell, temperature=300.0)
    tr = capi.mc_run(par, rep, [capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.1, labels=lab)], 1, 0, 0, 5)
    assert np.all(tr["accepted"] == -1) and np.array_equal(rep.positions, pos)  # the move fails, nothing changes
    lab[[3, 9]] = [0, 1]
    tr = capi.mc_run(par, rep, [capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=0.1, labels=lab)], 1, 0, 0, 40)
    assert set(tr["moved"].tolist()) == {3, 9}


def test_operation_geometry_like_the_reference_tests():
    """tests/operations/test_displacement_operations.py:19-112: |Sphere| = step, 0 < |Ball| < step, |Box| < step."""
    par = emt_params("Cu")
    pos, cell = fcc_cubic("Cu", 2)
    for op, check in [
        (capi.OP_SPHERE, lambda d: np.allclose(np.linalg.norm(d, axis=1), 0.2)),
        (capi.OP_BALL, lambda d: np.all((np.linalg.norm(d, axis=1) > 0) & (np.linalg.norm(d, axis=1) < 0.2))),
        (capi.OP_BOX, lambda d: np.all(np.abs(d) < 0.2)),
    ]:
        disp = []
        for a in range(40):
            rep = capi.Replica(positions=pos.copy(), cell=cell, temperature=1e7)  # everything accepted
            tr = capi.mc_run(par, rep, [capi.MoveSpec(capi.MOVE_DISPLACEMENT, op, step=0.2, labels=np.arange(32, dtype=np.int32))], 3, 0, a, 1)
            assert tr["accepted"][0] == 1
            i = tr["moved"][0]
            disp.append(rep.positions[i] - pos[i])
        assert check(np.array(disp))


def test_pt_swap_rule():
    lib = capi.lib()
    # (beta_a - beta_b)(E_a - E_b) with T_a < T_b, E_a < E_b is negative -> probabilistic
    x = (1 / (300 * units.kB) - 1 / (600 * units.kB)) * (1.0 - 2.0)
    assert lib.qo_pt_swap_accept(1.0, 2.0, 300.0, 600.0, math.exp(x) * 0.99) == 1
    assert lib.qo_pt_swap_accept(1.0, 2.0, 300.0, 600.0, min(0.999999, math.exp(x) * 1.01)) == (1 if math.exp(x) * 1.01 > 0.999999 else 0)
    assert lib.qo_pt_swap_accept(2.0, 1.0, 300.0, 600.0, 0.999999) == 1  # hot replica found the lower energy: always swap


def test_candidate_list_is_bit_identical_to_brute_force(monkeypatch):
    """Large orthorhombic cells (C5: 2048 atoms) visit a candidate list instead of every (atom, image) pair -- same order, same
    arithmetic, same cutoff test: energies, forces and sigma_1 must agree to the last bit with the brute-force loops."""
    import os

    from tests.helpers import cu_replicas

    pos, cell, n = cu_replicas(5, 1, seed=77, stdev=0.08)
    pos[0] += np.random.default_rng(5).integers(-2, 3, size=(n, 3)) * cell[0, 0]  # unwrapped coordinates
    par = emt_params("Cu")
    fast = capi.energy(par, pos[0], cell, (1, 1, 1), want_forces=True, want_sigma=True)
    monkeypatch.setenv("QO_NO_PAIRLIST", "1")
    brute = capi.energy(par, pos[0], cell, (1, 1, 1), want_forces=True, want_sigma=True)
    assert "QO_NO_PAIRLIST" in os.environ
    assert fast["n_pairs"] == brute["n_pairs"] and fast["energy"] == brute["energy"]
    assert np.array_equal(fast["forces"], brute["forces"]) and np.array_equal(fast["sigma1"], brute["sigma1"])
    slab = capi.energy(par, pos[0], cell, (1, 1, 0), want_forces=True)
    monkeypatch.delenv("QO_NO_PAIRLIST")
    slab_fast = capi.energy(par, pos[0], cell, (1, 1, 0), want_forces=True)
    assert slab["energy"] == slab_fast["energy"] and np.array_equal(slab["forces"], slab_fast["forces"])


def test_expm_of_symmetric_matrices_matches_scipy():
    """The deformation gradient of Anisotropic / ShapeDeformation is ``scipy.linalg.expm`` of a symmetric matrix
    (``operations/cell.py:96,141``); the oracle diagonalises by Jacobi rotations instead of scipy's Pade approximants."""
    from scipy.linalg import expm

    rng = np.random.default_rng(17)
    for scale in (1e-4, 0.01, 0.1, 1.0):
        for _ in range(20):
            c = rng.uniform(-scale, scale, 6)
            t = np.array([[c[0], c[3], c[4]], [c[3], c[1], c[5]], [c[4], c[5], c[2]]])
            np.testing.assert_allclose(capi.expm_sym3(t), expm(t), rtol=0, atol=2e-15 * max(1.0, float(np.exp(3 * scale))))
    np.testing.assert_array_equal(capi.expm_sym3(np.zeros((3, 3))), np.eye(3))
    d = np.diag([0.3, -0.2, 0.1])
    np.testing.assert_allclose(capi.expm_sym3(d), np.diag(np.exp([0.3, -0.2, 0.1])), rtol=0, atol=1e-16)


def test_emt_reproduces_the_numbers_of_ases_atomization_tutorial():
    """A second outside anchor for the ASE-owned arithmetic ("parity unpinned"): ASE's own documentation
    (tutorial "Atomization energy", `N2` with `EMT()`: d = 1.1 A) prints `Nitrogen atom energy: 5.10 eV`, `Nitrogen molecule energy:
    0.44 eV`, `Atomization energy: 9.76 eV`.  The restatements give 5.1 (an atom without neighbours: -E0), 0.440344 and 9.759656 --
    the printed digits -- with the cutoff built from the largest s0 of the WHOLE table, the gamma normalisations and the theta
    weight all taking part in the molecule's energy."""
    from oracle.potentials import emt_energy_forces, emt_params

    par = emt_params("N")
    box = np.eye(3) * 10.0
    atom = emt_energy_forces(np.zeros((1, 3)), box, (False,) * 3, par)["energy"]
    pos = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 1.1]])
    molecule = emt_energy_forces(pos, box, (False,) * 3, par)["energy"]
    assert f"{atom:.2f}" == "5.10" and f"{molecule:.2f}" == "0.44" and f"{2 * atom - molecule:.2f}" == "9.76"
    c = capi.energy(par, pos, box, (0, 0, 0), want_forces=True)
    assert abs(c["energy"] - molecule) <= 1e-13 and abs(c["forces"][0, 2] + c["forces"][1, 2]) <= 1e-13
    assert abs(molecule - 0.440343573035614) <= 1e-12  # (the value both restatements agree on, for regression)
