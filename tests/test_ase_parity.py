"""Parity gate against the GENUINE ase (pinned 3.26.0 by the reference, uv.lock:6-7).

``ase`` is a third-party dependency that is neither vendored in the reference nor installable in the build container, so the
ASE-owned arithmetic of the hot path -- ``ase.calculators.emt.EMT``, ``ase.calculators.lj.LennardJones``, the neighbour image
enumeration behind them (call sites: src/quansino/mc/criteria.py:105,133,163,241 and integrators/displacement.py:63,77) -- is
restated in ``oracle/potentials.py`` / ``oracle/qmc_oracle.c``.  Wherever ``ase`` IS importable this module pins those
restatements: energies, forces and per-atom energies against the real calculators to 1e-12, and the golden traces regenerated
through real ASE (``oracle/gen_golden.py --real-ase``).  Without ``ase`` the module skips (and DESIGN.md says "parity unpinned"
for the ASE-owned part).
"""

import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ase = pytest.importorskip("ase", reason="the genuine ase is not installed: ASE-owned arithmetic stays unpinned here")

from oracle import capi  # noqa: E402
from oracle.potentials import LJParams, emt_energy_forces, emt_params, lj_energy_forces  # noqa: E402
from quansino_b200.systems import fcc_cubic, octahedron, rattle  # noqa: E402

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"


def _configurations():
    """(name, symbol, positions, cell, pbc, potential) -- the golden inputs plus the shapes of BASELINE.json's configurations."""
    out = []
    for f in sorted(GOLDEN.glob("*.npz")):
        g = np.load(f, allow_pickle=True)
        if "positions0" not in g.files:
            continue
        lj = "sigma" in g.files
        out.append((f.stem, "Pt" if lj else "Cu", g["positions0"], g["cell"], (not lj,) * 3,
                    ("lj", float(g["sigma"]), float(g["epsilon"])) if lj else ("emt",)))  # fmt: skip
    for repeat, stdev in ((2, 0.08), (3, 0.05), (5, 0.05)):
        pos, cell = fcc_cubic("Cu", repeat)
        out.append((f"cu{repeat}", "Cu", rattle(pos, stdev, 100 + repeat), cell, (True,) * 3, ("emt",)))
    pos, cell = fcc_cubic("Cu", 3)
    out.append(("cu3_slab", "Cu", rattle(pos, 0.05, 9), cell, (True, True, False), ("emt",)))
    out.append(("cu3_unwrapped", "Cu", rattle(pos, 0.05, 10) + np.random.default_rng(1).integers(-3, 4, (len(pos), 3)) * cell[0, 0],
                cell, (True,) * 3, ("emt",)))  # fmt: skip
    pos, cell = octahedron("Pt", 4, 10.0)
    out.append(("pt44", "Pt", rattle(pos, 0.02, 99), cell, (False,) * 3, ("lj", 2.5, 0.1)))
    return out


@pytest.mark.parametrize("case", _configurations(), ids=lambda c: c[0])
def test_restated_potentials_match_real_ase(case):
    from ase import Atoms
    from ase.calculators.emt import EMT
    from ase.calculators.lj import LennardJones

    name, symbol, pos, cell, pbc, pot = case
    if pot[0] == "emt":
        calc, par = EMT(), emt_params(symbol)
        res = emt_energy_forces(pos, cell, pbc, par)
    else:
        calc = LennardJones(sigma=pot[1], epsilon=pot[2])
        par = LJParams(sigma=pot[1], epsilon=pot[2])  # rc = None -> 3 sigma, smooth = False: ASE's defaults
        res = lj_energy_forces(pos, cell, pbc, par)
    e_np, f_np = res["energy"], res["forces"]
    atoms = Atoms(f"{symbol}{len(pos)}", positions=pos, cell=cell, pbc=pbc, calculator=calc)
    e_ase, f_ase = atoms.get_potential_energy(), atoms.get_forces()
    ref = capi.energy(par, pos, cell, tuple(int(b) for b in pbc), want_forces=True)
    scale = max(abs(e_ase), 1.0)
    assert abs(e_np - e_ase) <= 1e-12 * scale, f"{name}: numpy restatement {e_np!r} vs ase {e_ase!r}"
    assert abs(ref["energy"] - e_ase) <= 1e-12 * scale, f"{name}: C oracle {ref['energy']!r} vs ase {e_ase!r}"
    np.testing.assert_allclose(f_np, f_ase, rtol=0, atol=1e-12 * max(np.abs(f_ase).max(), 1.0))
    np.testing.assert_allclose(ref["forces"], f_ase, rtol=0, atol=1e-12 * max(np.abs(f_ase).max(), 1.0))
    if "energies" in calc.results:  # what AdaptiveForceBias reads (mc/fbmc.py:329)
        assert abs(np.sum(calc.results["energies"]) - e_ase) <= 1e-12 * scale


def test_ase_units_match_the_restatement():
    import ase.units

    from oracle import units

    for name in ("kB", "fs", "bar", "Bohr"):
        if hasattr(units, name):
            assert getattr(units, name) == getattr(ase.units, name), name


@pytest.mark.skipif(not Path("/root/reference/src/quansino").exists(), reason="the reference sources are not on this machine")
def test_golden_traces_regenerate_through_real_ase():
    """The genuine quansino sources over the GENUINE ase reproduce every committed golden trace (which were generated over the
    shim): integer traces identical, energies to 1e-10."""
    res = subprocess.run([sys.executable, str(ROOT / "oracle" / "gen_golden.py"), "--real-ase"], capture_output=True, text=True, timeout=1800)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
