#!/usr/bin/env python
"""Generate golden traces by running the GENUINE quansino sources.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Runs the unmodified reference (``/root/reference/src/quansino``, v0.1.3) in this container:

* ``ase`` is absent, so ``oracle/ase_shim`` stands in for it (``Atoms``, ``Cell``, units, calculator cache) and its
  ``EMT`` / ``LennardJones`` evaluate the numpy restatement ``oracle/potentials.py``;
* ``importlib.metadata.version("quansino")`` is answered with "0.1.3" (the tree is not pip-installed);
* the drivers' ``rng`` is replaced by ``oracle.philox.InjectedStream`` -- quansino only stores and calls it
  (``src/quansino/mc/contexts.py:49-52``) -- so the draws are the Philox slots the kernels use.

Everything else -- move selection, proposals, criteria, save / revert, label bookkeeping, Verlet -- is quansino's own
code.  The traces pin the C oracle (tests/test_golden.py) and the CUDA path (tests/test_gpu_golden.py).

Usage (in the build container only; the GPU box has no /root/reference):  python oracle/gen_golden.py [case ...]

``--real-ase`` (wherever the genuine ``ase`` is importable): the same runs over REAL ASE instead of ``oracle/ase_shim``; nothing
is written -- every trace is compared with the committed golden file (integer traces identical, floats to 1e-10 relative /
1e-12 A) and the script exits non-zero on any difference.  ``tests/test_ase_parity.py`` runs it when ``ase`` is present.  This
is the check that pins the ASE-owned arithmetic (EMT, Lennard-Jones, image enumeration, ``euler_rotate``) that the shim and
``oracle/potentials.py`` restate.
"""

from __future__ import annotations

import importlib.metadata as _md
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REAL_ASE = "--real-ase" in sys.argv
if REAL_ASE:
    sys.argv.remove("--real-ase")
    sys.path[:0] = [str(ROOT), "/root/reference/src"]  # the genuine ase from site-packages
else:
    sys.path[:0] = [str(ROOT), str(ROOT / "oracle" / "ase_shim"), "/root/reference/src"]

_orig_version = _md.version
_md.version = lambda name: "0.1.3" if name == "quansino" else _orig_version(name)

from ase import Atoms  # noqa: E402  (the shim)
from ase.calculators.emt import EMT  # noqa: E402
from ase.calculators.lj import LennardJones  # noqa: E402
from quansino.integrators.displacement import Verlet  # noqa: E402
from quansino.mc.canonical import Canonical, HamiltonianCanonical  # noqa: E402
from quansino.mc.criteria import CanonicalCriteria, GrandCanonicalCriteria, IsobaricCriteria, IsotensionCriteria  # noqa: E402
from quansino.mc.fbmc import AdaptiveForceBias, ForceBias  # noqa: E402
from quansino.moves.composite import CompositeMove  # noqa: E402
from quansino.mc.gcmc import GrandCanonical  # noqa: E402
from quansino.mc.isobaric import Isobaric  # noqa: E402
from quansino.mc.isotension import Isotension  # noqa: E402
from quansino.moves.cell import CellMove  # noqa: E402
from quansino.moves.displacement import DisplacementMove, HamiltonianDisplacementMove  # noqa: E402
from quansino.moves.exchange import CompositeExchangeMove, ExchangeMove  # noqa: E402
from quansino.operations.cell import AnisotropicDeformation, IsotropicDeformation, ShapeDeformation  # noqa: E402
from quansino.operations.displacement import Ball, Box, Rotation, Sphere, Translation, TranslationRotation  # noqa: E402

from oracle.philox import InjectedStream  # noqa: E402
from oracle.units import bar  # noqa: E402
from quansino_b200.systems import fcc_cubic, octahedron, rattle  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


def inject(mc, seed, replica):
    stream = InjectedStream(seed, replica)
    mc._rng = stream
    mc.context.rng = stream
    return stream


def run_steps(mc, steps):
    """Drive ``MonteCarlo.step`` exactly like ``Driver.irun`` (mc/driver.py:91-119) and record every attempt."""
    mc.validate_simulation()
    rec = {"move": [], "accepted": [], "energy": [], "n_atoms": []}
    names = list(mc.moves)

    def record():
        name, acc = mc.move_history[-1]
        rec["move"].append(names.index(name))
        rec["accepted"].append(-1 if acc is None else int(bool(acc)))
        rec["energy"].append(float(mc.context.last_potential_energy))
        rec["n_atoms"].append(len(mc.atoms))

    for _ in range(steps):
        seen = 0
        for _name in mc.step():
            if hasattr(mc.context.rng, "mark_cycle"):
                mc.context.rng.mark_cycle()
            if len(mc.move_history) > seen:
                record()
                seen = len(mc.move_history)
        if len(mc.move_history) > seen:
            record()
        mc.step_count += 1
    return {
        "move": np.array(rec["move"], dtype=np.int8),
        "accepted": np.array(rec["accepted"], dtype=np.int8),
        "energy": np.array(rec["energy"]),
        "n_atoms": np.array(rec["n_atoms"], dtype=np.int32),
    }


def cu_atoms(repeat, stdev, seed):
    pos, cell = fcc_cubic("Cu", repeat)
    pos = rattle(pos, stdev, seed)
    return Atoms(f"Cu{len(pos)}", positions=pos, cell=cell, pbc=True, calculator=EMT()), pos, cell


def case_canonical(opname, op, seed, steps):
    atoms, pos0, cell = cu_atoms(3, 0.05, 4242)
    n = len(atoms)
    mc = Canonical(atoms, temperature=300.0, default_displacement_move=DisplacementMove(np.arange(n), op(0.1)))
    inject(mc, seed, 3)
    tr = run_steps(mc, steps)
    return dict(
        kind="canonical", op=opname, step_size=0.1, seed=seed, replica=3, temperature=300.0, positions0=pos0, cell=cell,
        positions=atoms.positions.copy(), n_calculations=getattr(atoms.calc, "n_calculations", -1), **tr,
    )  # fmt: skip


def case_composite(variant, seed, n_attempts):
    """CompositeDisplacementMove (moves/displacement.py:363-455) built with the reference's own `*` and `+`."""
    atoms, pos0, cell = cu_atoms(3, 0.05, 5151)
    n = len(atoms)
    labels = np.arange(n)
    if variant == "mul":  # a plain move and a triple move, selected with equal probability
        mc = Canonical(atoms, temperature=300.0, max_cycles=n_attempts, default_displacement_move=DisplacementMove(labels, Ball(0.1)))
        mc.add_move(DisplacementMove(labels, Ball(0.06)) * 3, criteria=CanonicalCriteria(), name="triple", probability=1.0)
        spec = [[("Ball", 0.1, 1)], [("Ball", 0.06, 3)]]
    elif variant == "add":  # three different operations in one composite
        mc = Canonical(atoms, temperature=300.0, max_cycles=n_attempts)
        mc.add_move(DisplacementMove(labels, Ball(0.08)) + DisplacementMove(labels, Box(0.04)) + DisplacementMove(labels, Sphere(0.05)),
                    criteria=CanonicalCriteria(), name="mixed")  # fmt: skip
        spec = [[("Ball", 0.08, 1), ("Box", 0.04, 1), ("Sphere", 0.05, 1)]]
    elif variant == "mul24":  # more sub-moves than the first block range holds (sub-moves 17 and up draw from blocks 0x8000 ...)
        mc = Canonical(atoms, temperature=300.0, max_cycles=n_attempts, default_displacement_move=DisplacementMove(labels, Ball(0.1)))
        mc.add_move(DisplacementMove(labels, Ball(0.03)) * 24, criteria=CanonicalCriteria(), name="many", probability=1.0)
        spec = [[("Ball", 0.1, 1)], [("Ball", 0.03, 24)]]
    elif variant == "exhaust":  # only two movable atoms: the third and fourth sub-moves find no candidate
        labels = np.full(n, -1)
        labels[7] = 0
        labels[40] = 1
        mc = Canonical(atoms, temperature=300.0, max_cycles=n_attempts)
        mc.add_move(DisplacementMove(labels, Ball(0.1)) * 4, criteria=CanonicalCriteria(), name="four")
        spec = [[("Ball", 0.1, 4)]]
    else:
        raise ValueError(variant)
    inject(mc, seed, 4)
    tr = run_steps(mc, 1)
    flat = np.array([(i, ["Ball", "Box", "Sphere"].index(o), s, r) for i, chain in enumerate(spec) for (o, s, r) in chain], dtype=float)
    return dict(
        kind="composite", variant=variant, seed=seed, replica=4, temperature=300.0, positions0=pos0, cell=cell, labels=labels,
        spec=flat, probabilities=np.array([mc.moves[k].probability for k in mc.moves]), positions=atoms.positions.copy(), **tr,
    )  # fmt: skip


def case_groups(seed, n_attempts):
    """Label groups (moves/displacement.py:164-167): every atom carrying the selected label moves by the same translation.
    Dimers, one group of five, single atoms, frozen atoms (label -1); label values are neither dense nor ordered by atom."""
    atoms, pos0, cell = cu_atoms(3, 0.05, 7171)
    n = len(atoms)
    labels = np.full(n, -1)
    labels[0:40] = 100 + np.repeat(np.arange(20), 2)  # dimers (0,1), (2,3), ...
    labels[[50, 61, 72, 83, 94]] = 7  # one group of five, with the SMALLEST label
    labels[100:108] = 300 - np.arange(8)  # single atoms, labels decreasing with the atom index
    mc = Canonical(atoms, temperature=300.0, max_cycles=n_attempts, default_displacement_move=DisplacementMove(labels, Ball(0.06)))
    inject(mc, seed, 7)
    tr = run_steps(mc, 1)
    return dict(
        kind="groups", seed=seed, replica=7, temperature=300.0, step_size=0.06, positions0=pos0, cell=cell, labels=labels,
        positions=atoms.positions.copy(), **tr,
    )  # fmt: skip


def case_rotation(seed, n_attempts, op_name="Rotation"):
    """Rotation (operations/displacement.py:178-200) of label groups -- trimers and one group of six -- next to rigid translations
    of the same groups; ``op_name`` "Translation" / "TranslationRotation" (:156-175, :203-227) relocates (and turns) the group
    instead: centroid to a uniform point of the cell."""
    atoms, pos0, cell = cu_atoms(3, 0.05, 9191)
    n = len(atoms)
    labels = np.full(n, -1)
    labels[0:36] = np.repeat(np.arange(12), 3) * 2 + 11  # trimers (0,1,2), (3,4,5), ...
    labels[[40, 51, 62, 73, 84, 95]] = 4
    mc = Canonical(atoms, temperature=4000.0, max_cycles=n_attempts, default_displacement_move=DisplacementMove(labels, Ball(0.05)))
    op = {"Rotation": Rotation, "Translation": Translation, "TranslationRotation": TranslationRotation}[op_name]()
    mc.add_move(DisplacementMove(labels, op), name="rotation", probability=1.0)
    inject(mc, seed, 9)
    tr = run_steps(mc, 1)
    return dict(
        kind="rotation", op=op_name, seed=seed, replica=9, temperature=4000.0, step_size=0.05, positions0=pos0, cell=cell, labels=labels,
        positions=atoms.positions.copy(), **tr,
    )  # fmt: skip


def case_fbmc(seed, steps, delta, temperature, adaptive=None):
    """ForceBias (mc/fbmc.py:202-241): every atom moves every step, zeta drawn by rejection against the force-biased trial
    probability; no acceptance test.  ``adaptive = (min_delta, max_delta, scheme, update_function, reference_variance)`` runs
    AdaptiveForceBias (mc/fbmc.py:285-488) instead: EMT has no committee, so every step warns and falls back to the reference
    variance -- delta = min + (max - min) * update(reference_variance)."""
    import warnings

    atoms, pos0, cell = cu_atoms(3, 0.05, 1212)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # "No FixCom constraint found"
        if adaptive:
            fb = AdaptiveForceBias(atoms, adaptive[0], adaptive[1], temperature=temperature, scheme=adaptive[2],
                                   update_function=adaptive[3], reference_variance=adaptive[4], seed=1)  # fmt: skip
        else:
            fb = ForceBias(atoms, delta=delta, temperature=temperature, seed=1)
    fb._rng = InjectedStream(seed, 10, mode="fbmc")
    traj, energies, gmax, deltas = [], [], [], []
    for _ in range(steps):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # "No committee forces available"
            fb.step()
        delta = float(np.max(fb.delta))
        assert np.all(np.asarray(fb.delta) == delta)
        deltas.append(delta)
        traj.append(atoms.positions.copy())
        energies.append(float(atoms.get_potential_energy()))
        gmax.append(float(np.max(np.abs(fb.gamma))))
    return dict(kind="fbmc", seed=seed, replica=10, temperature=temperature, delta=delta, delta_steps=np.array(deltas), positions0=pos0, cell=cell,
                trajectory=np.array(traj), energy=np.array(energies), gamma_max=np.array(gmax), positions=atoms.positions.copy(),
                accepted=np.ones(steps, np.int8), move=np.zeros(steps, np.int8), n_atoms=np.full(steps, len(atoms), np.int32))  # fmt: skip


def case_composite_operation(seed, n_attempts):
    """CompositeOperation (operations/composite.py:22-109): `Ball(0.06) + Box(0.03) + Sphere(0.02)` and `Box(0.04) * 2` built with
    the reference's own operator algebra; the displacement is the sum of the parts."""
    atoms, pos0, cell = cu_atoms(3, 0.05, 2323)
    n = len(atoms)
    mc = Canonical(atoms, temperature=400.0, max_cycles=n_attempts,
                   default_displacement_move=DisplacementMove(np.arange(n), Ball(0.06) + Box(0.03) + Sphere(0.02)))  # fmt: skip
    mc.add_move(DisplacementMove(np.arange(n), Box(0.04) * 2), name="double_box", probability=1.0)
    inject(mc, seed, 11)
    tr = run_steps(mc, 1)
    return dict(kind="composite_operation", seed=seed, replica=11, temperature=400.0, positions0=pos0, cell=cell,
                positions=atoms.positions.copy(), **tr)  # fmt: skip


def case_hybrid(seed, n_attempts, k):
    """The docs' hybrid NPT move (examples/orb-v3_isobaric_bulk_cu.py:45-48): `DisplacementMove * k + CellMove` is a plain
    CompositeMove -- k independent displacements (labels may repeat), one cell deformation, ONE IsobaricCriteria test -- next
    to a plain displacement move; plus a plain CompositeMove of two displacement moves under CanonicalCriteria."""
    atoms, pos0, cell = cu_atoms(3, 0.05, 6161)
    n = len(atoms)
    labels = np.arange(n)
    pressure = 2000.0 * bar
    mc = Isobaric(atoms, temperature=500.0, pressure=pressure, max_cycles=n_attempts,
                  default_displacement_move=DisplacementMove(labels, Ball(0.1)))  # fmt: skip
    hybrid = DisplacementMove(labels, Ball(0.05)) * k + CellMove(IsotropicDeformation(0.01))
    assert type(hybrid) is CompositeMove and len(hybrid) == k + 1
    mc.add_move(hybrid, criteria=IsobaricCriteria(), name="hybrid", probability=1.0)
    pair = CompositeMove([DisplacementMove(labels, Box(0.04)), DisplacementMove(labels, Box(0.04))])
    mc.add_move(pair, criteria=CanonicalCriteria(), name="pair", probability=0.5)
    mc.moves["default_displacement_move"].probability = 1.0
    inject(mc, seed, 6)
    tr = run_steps(mc, 1)
    # (named move, type 0 = displacement / 1 = cell, operation, step, repeat, chain bits)
    spec = np.array([(0, 0, 0, 0.1, 1, 0), (1, 0, 0, 0.05, k, 3), (1, 1, 4, 0.01, 1, 2), (2, 0, 1, 0.04, 2, 2)], dtype=float)
    return dict(
        kind="hybrid", seed=seed, replica=6, temperature=500.0, pressure=pressure, positions0=pos0, cell0=cell, labels=labels, spec=spec,
        probabilities=np.array([mc.moves[q].probability for q in mc.moves]), positions=atoms.positions.copy(), cell=atoms.cell.array.copy(), **tr,
    )  # fmt: skip


def case_forced(seed, steps):
    """Forced moves (minimum_count, mc/core.py:331-342): two cell moves and three Box displacements are guaranteed per step of 40
    cycles; the remaining cycles choose between the default displacement move and the (rarely selected) others."""
    atoms, pos0, cell = cu_atoms(3, 0.05, 8181)
    n = len(atoms)
    pressure = 1500.0 * bar
    mc = Isobaric(atoms, temperature=400.0, pressure=pressure, max_cycles=40, default_displacement_move=DisplacementMove(np.arange(n), Ball(0.1)),
                  default_cell_move=CellMove(IsotropicDeformation(0.01)))  # fmt: skip
    mc.moves["default_displacement_move"].probability = 1.0
    mc.moves["default_cell_move"].probability = 0.02
    mc.moves["default_cell_move"].minimum_count = 2
    mc.add_move(DisplacementMove(np.arange(n), Box(0.05)), name="box", probability=0.05, minimum_count=3)
    inject(mc, seed, 8)
    tr = run_steps(mc, steps)
    return dict(
        kind="forced", seed=seed, replica=8, temperature=400.0, pressure=pressure, max_cycles=40, positions0=pos0, cell0=cell,
        probabilities=np.array([mc.moves[q].probability for q in mc.moves]), minimum_counts=np.array([0, 2, 3]),
        positions=atoms.positions.copy(), cell=atoms.cell.array.copy(), **tr,
    )  # fmt: skip


def case_isobaric(seed, steps, p_cell, mask=None):
    atoms, pos0, cell = cu_atoms(3, 0.05, 777)
    n = len(atoms)
    pressure = 1000.0 * bar
    mc = Isobaric(
        atoms, temperature=300.0, pressure=pressure, default_displacement_move=DisplacementMove(np.arange(n), Ball(0.1)),
        default_cell_move=CellMove(IsotropicDeformation(0.05, mask=mask)),
    )  # fmt: skip
    default_probabilities = np.array([mc.moves["default_displacement_move"].probability, mc.moves["default_cell_move"].probability])
    if p_cell is not None:
        mc.moves["default_cell_move"].probability = p_cell
        mc.moves["default_displacement_move"].probability = 1 - p_cell
    probabilities = np.array([mc.moves["default_displacement_move"].probability, mc.moves["default_cell_move"].probability])
    inject(mc, seed, 1)
    tr = run_steps(mc, steps)
    return dict(
        kind="isobaric", seed=seed, replica=1, temperature=300.0, pressure=pressure, step_size=0.1, max_value=0.05,
        default_probabilities=default_probabilities, probabilities=probabilities, positions0=pos0, cell0=cell,
        positions=atoms.positions.copy(), cell=atoms.cell.array.copy(),
        axes=7 if mask is None else int(sum(int(mask[d, d]) << d for d in range(3))), **tr,
    )  # fmt: skip


def case_general_cell(seed, n_attempts, op_name, max_value, p_cell, stress=None, shear=None, mask=None):
    """Cell moves that leave GENERAL (triclinic) cells behind (operations/cell.py:56-141: AnisotropicDeformation, ShapeDeformation;
    IsotropicDeformation of an already sheared cell), judged by IsobaricCriteria (mc/criteria.py:143-172) or, with `stress`, by
    IsotensionCriteria through the Isotension driver (mc/criteria.py:175-219, mc/isotension.py:27-121); displacement moves in
    between see the deformed cell."""
    atoms, pos0, cell = cu_atoms(3, 0.05, 24680 + seed)
    if shear is not None:  # start from a sheared cell: cell' = S @ cell, atoms carried along
        atoms.set_cell(np.asarray(shear) @ cell, scale_atoms=True)
        pos0, cell = atoms.positions.copy(), atoms.cell.array.copy()
    n = len(atoms)
    pressure = 1000.0 * bar
    op = {"Anisotropic": AnisotropicDeformation, "Shape": ShapeDeformation, "Isotropic": IsotropicDeformation}[op_name](max_value, mask=mask)
    common = dict(temperature=600.0, pressure=pressure, max_cycles=n_attempts,
                  default_displacement_move=DisplacementMove(np.arange(n), Ball(0.1)), default_cell_move=CellMove(op))  # fmt: skip
    mc = Isobaric(atoms, **common) if stress is None else Isotension(atoms, external_stress=np.asarray(stress, dtype=float), **common)
    assert isinstance(mc.moves["default_cell_move"].criteria, IsobaricCriteria if stress is None else IsotensionCriteria)
    mc.moves["default_cell_move"].probability = p_cell
    mc.moves["default_displacement_move"].probability = 1 - p_cell
    inject(mc, seed, 5)
    tr = run_steps(mc, 1)
    m9 = np.ones((3, 3), dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
    return dict(
        kind="general_cell", seed=seed, replica=5, temperature=600.0, pressure=pressure, step_size=0.1, max_value=max_value, op=op_name,
        p_cell=p_cell, isotension=int(stress is not None), external_stress=np.zeros((3, 3)) if stress is None else np.asarray(stress, dtype=float),
        mask=m9, positions0=pos0, cell0=cell, positions=atoms.positions.copy(), cell=atoms.cell.array.copy(), **tr,
    )  # fmt: skip


def case_gc_lj(seed, n_attempts, mu):
    pos0, cell = octahedron("Pt", 4, 10.0)
    pos0 = rattle(pos0, 0.02, 99)
    n = len(pos0)
    atoms = Atoms(f"Pt{n}", positions=pos0, cell=cell, pbc=False, calculator=LennardJones(sigma=2.5, epsilon=0.1))
    disp = DisplacementMove(np.arange(n))
    exch = ExchangeMove(np.full(n, -1))
    mc = GrandCanonical(
        atoms, exchange_atoms=Atoms("Pt"), temperature=300.0, chemical_potential=mu, number_of_exchange_particles=0,
        max_cycles=n_attempts, default_displacement_move=disp, default_exchange_move=exch,
    )  # fmt: skip
    mc.moves["default_displacement_move"].probability = 0.5
    mc.moves["default_exchange_move"].probability = 0.5
    inject(mc, seed, 2)
    tr = run_steps(mc, 1)
    return dict(
        kind="gc_lj", seed=seed, replica=2, temperature=300.0, chemical_potential=mu, sigma=2.5, epsilon=0.1, step_size=0.1,
        accessible_volume=float(mc.accessible_volume), exchange_mass=float(mc.exchange_atoms.get_masses().sum()),
        positions0=pos0, cell=cell, positions=atoms.positions.copy(), labels_displacement=np.asarray(disp.labels),
        labels_exchange=np.asarray(exch.labels), n_exchange=int(mc.number_of_exchange_particles), **tr,
    )  # fmt: skip


def case_gc_molecule(seed, n_attempts, mu):
    """Grand-canonical insertion / deletion of a DIMER (two atoms sharing one label) placed by TranslationRotation, next to
    displacement moves whose inserted molecules become label groups (moves/exchange.py:88-176, moves/displacement.py:226-253,
    operations/displacement.py:203-227).  NOT lowered to the device yet: the trace is the fixture the next
    step starts from; tests/test_golden.py checks its bookkeeping invariants."""
    pos0, cell = octahedron("Pt", 4, 10.0)
    pos0 = rattle(pos0, 0.02, 98)
    n = len(pos0)
    atoms = Atoms(f"Pt{n}", positions=pos0, cell=cell, pbc=False, calculator=LennardJones(sigma=2.5, epsilon=0.1))
    disp = DisplacementMove(np.arange(n))
    exch = ExchangeMove(np.full(n, -1), TranslationRotation())
    dimer = Atoms("Pt2", positions=[[0.0, 0.0, 0.0], [0.0, 0.0, 2.8]])
    mc = GrandCanonical(
        atoms, exchange_atoms=dimer, temperature=300.0, chemical_potential=mu, number_of_exchange_particles=0,
        max_cycles=n_attempts, default_displacement_move=disp, default_exchange_move=exch,
    )  # fmt: skip
    mc.moves["default_displacement_move"].probability = 0.5
    mc.moves["default_exchange_move"].probability = 0.5
    inject(mc, seed, 3)
    tr = run_steps(mc, 1)
    return dict(
        kind="gc_molecule", seed=seed, replica=3, temperature=300.0, chemical_potential=mu, sigma=2.5, epsilon=0.1, step_size=0.1,
        accessible_volume=float(mc.accessible_volume), exchange_mass=float(mc.exchange_atoms.get_masses().sum()),
        exchange_positions=dimer.positions.copy(), positions0=pos0, cell=cell, positions=atoms.positions.copy(),
        labels_displacement=np.asarray(disp.labels), labels_exchange=np.asarray(exch.labels),
        n_exchange=int(mc.number_of_exchange_particles), **tr,
    )  # fmt: skip


def case_gc_composite(seed, n_attempts, mu, temperature, k=2):
    """``CompositeExchangeMove`` of k single-atom ``ExchangeMove`` objects with equal labels (moves/exchange.py:237-316): one coin,
    k insertions with their own Translation draws or k label deletions without replacement, ONE GrandCanonicalCriteria evaluation
    with particle_delta = +-k (mc/criteria.py:225-277); after an accepted insertion every move gives the k new atoms one label
    (moves/displacement.py:244-249), so that they are displaced and deleted together from then on."""
    pos0, cell = octahedron("Pt", 4, 10.0)
    pos0 = rattle(pos0, 0.02, 97)
    n = len(pos0)
    atoms = Atoms(f"Pt{n}", positions=pos0, cell=cell, pbc=False, calculator=LennardJones(sigma=2.5, epsilon=0.1))
    disp = DisplacementMove(np.arange(n))
    subs = [ExchangeMove(np.full(n, -1)) for _ in range(k)]
    exch = subs[0]
    for sub in subs[1:]:
        exch = exch + sub
    assert isinstance(exch, CompositeExchangeMove) and len(exch) == k
    mc = GrandCanonical(
        atoms, exchange_atoms=Atoms("Pt"), temperature=temperature, chemical_potential=mu, number_of_exchange_particles=0,
        max_cycles=n_attempts, default_displacement_move=disp,
    )  # fmt: skip
    mc.add_move(exch, criteria=GrandCanonicalCriteria(), name="default_exchange_move")
    mc.moves["default_displacement_move"].probability = 0.5
    mc.moves["default_exchange_move"].probability = 0.5
    inject(mc, seed, 4)
    tr = run_steps(mc, 1)
    for sub in subs[1:]:
        assert np.array_equal(sub.labels, subs[0].labels)
    return dict(
        kind="gc_composite", seed=seed, replica=4, temperature=temperature, chemical_potential=mu, sigma=2.5, epsilon=0.1, step_size=0.1,
        k=k, bias=float(exch.bias_towards_insert), accessible_volume=float(mc.accessible_volume),
        exchange_mass=float(mc.exchange_atoms.get_masses().sum()), positions0=pos0, cell=cell, positions=atoms.positions.copy(),
        labels_displacement=np.asarray(disp.labels), labels_exchange=np.asarray(subs[0].labels),
        n_exchange=int(mc.number_of_exchange_particles), **tr,
    )  # fmt: skip


def case_gc_emt(seed, n_attempts, mu, temperature, shear=None):
    atoms, pos0, cell = cu_atoms(3, 0.05, 31337)
    if shear is not None:  # a general (triclinic) cell: cell' = S @ cell, atoms carried along
        atoms.set_cell(np.asarray(shear) @ cell, scale_atoms=True)
        pos0, cell = atoms.positions.copy(), atoms.cell.array.copy()
    n = len(atoms)
    disp = DisplacementMove(np.arange(n))
    exch = ExchangeMove(np.arange(n))
    mc = GrandCanonical(
        atoms, exchange_atoms=Atoms("Cu"), temperature=temperature, chemical_potential=mu, number_of_exchange_particles=n,
        max_cycles=n_attempts, default_displacement_move=disp, default_exchange_move=exch,
    )  # fmt: skip
    mc.moves["default_displacement_move"].probability = 0.3
    mc.moves["default_exchange_move"].probability = 0.7
    inject(mc, seed, 0)
    tr = run_steps(mc, 1)
    return dict(
        kind="gc_emt", seed=seed, replica=0, temperature=temperature, chemical_potential=mu, step_size=0.1,
        accessible_volume=float(mc.accessible_volume), exchange_mass=float(mc.exchange_atoms.get_masses().sum()),
        positions0=pos0, cell=cell, positions=atoms.positions.copy(), labels_displacement=np.asarray(disp.labels),
        labels_exchange=np.asarray(exch.labels), n_exchange=int(mc.number_of_exchange_particles), n_exchange0=n, **tr,
    )  # fmt: skip


def case_hmc(seed, n_attempts, dt_fs, n_steps, temperature, shear=None):
    atoms, pos0, cell = cu_atoms(3, 0.03, 2718)
    if shear is not None:  # a general (triclinic) cell
        atoms.set_cell(np.asarray(shear) @ cell, scale_atoms=True)
        pos0, cell = atoms.positions.copy(), atoms.cell.array.copy()
    mc = HamiltonianCanonical(
        atoms, temperature=temperature, max_cycles=1,
        default_displacement_move=HamiltonianDisplacementMove(operation=Verlet(dt=dt_fs, max_steps=n_steps)),
    )  # fmt: skip
    inject(mc, seed, 5)
    tr = run_steps(mc, n_attempts)
    return dict(
        kind="hmc", seed=seed, replica=5, temperature=temperature, dt_fs=dt_fs, n_steps=n_steps, positions0=pos0, cell=cell,
        positions=atoms.positions.copy(), momenta=atoms.get_momenta(), last_kinetic=float(mc.context.last_kinetic_energy), **tr,
    )  # fmt: skip


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    cases = {
        "canonical_ball": lambda: case_canonical("Ball", Ball, 1001, 2),
        "canonical_box": lambda: case_canonical("Box", Box, 1002, 1),
        "canonical_sphere": lambda: case_canonical("Sphere", Sphere, 1003, 1),
        "isobaric_default": lambda: case_isobaric(2001, 1, None),
        "isobaric_pcell": lambda: case_isobaric(2002, 1, 0.2),
        "isobaric_masked": lambda: case_isobaric(2003, 1, 0.25, mask=np.diag([True, False, True])),
        "npt_anisotropic": lambda: case_general_cell(2101, 160, "Anisotropic", 0.008, 0.25),
        "nst_shape": lambda: case_general_cell(2102, 160, "Shape", 0.012, 0.25, stress=np.array([[2e-3, 1e-3, 0.0], [1e-3, -1e-3, 5e-4], [0.0, 5e-4, 1e-3]])),
        "nst_anisotropic_masked": lambda: case_general_cell(
            2103, 160, "Anisotropic", 0.008, 0.25, stress=np.diag([1e-3, 2e-3, -1e-3]),
            mask=np.array([[True, True, False], [True, True, False], [False, False, True]])),
        "npt_isotropic_sheared": lambda: case_general_cell(2104, 120, "Isotropic", 0.02, 0.2, shear=[[1.0, 0.0, 0.0], [0.12, 1.0, 0.0], [-0.08, 0.1, 1.0]]),
        "gc_lj": lambda: case_gc_lj(3001, 700, -0.5),
        "gc_emt": lambda: case_gc_emt(3002, 150, -3.4, 3000.0),
        "gc_molecule": lambda: case_gc_molecule(3003, 600, -0.6),
        "gc_emt_sheared": lambda: case_gc_emt(3005, 200, -3.8, 3000.0, shear=[[1.0, 0.02, 0.0], [0.05, 1.0, -0.02], [-0.03, 0.04, 1.0]]),
        "gc_composite": lambda: case_gc_composite(3004, 600, -3.3, 2000.0),
        "hmc": lambda: case_hmc(4001, 3, 2.0, 12, 800.0),
        "hmc_sheared": lambda: case_hmc(4002, 3, 2.0, 12, 800.0, shear=[[1.0, 0.02, 0.0], [0.05, 1.0, -0.02], [-0.03, 0.04, 1.0]]),
        "composite_mul": lambda: case_composite("mul", 5001, 120),
        "composite_add": lambda: case_composite("add", 5002, 80),
        "composite_exhaust": lambda: case_composite("exhaust", 5003, 40),
        "composite_mul24": lambda: case_composite("mul24", 5005, 60),
        "composite_hybrid": lambda: case_hybrid(5004, 90, 3),
        "label_groups": lambda: case_groups(6001, 100),
        "forced_moves": lambda: case_forced(7001, 3),
        "group_rotation": lambda: case_rotation(8001, 100),
        "group_translation": lambda: case_rotation(8002, 60, "Translation"),
        "group_translation_rotation": lambda: case_rotation(8003, 100, "TranslationRotation"),
        "force_bias": lambda: case_fbmc(9001, 4, 0.1, 600.0),
        "composite_operation": lambda: case_composite_operation(9501, 90),
        "adaptive_force_bias": lambda: case_fbmc(9002, 3, None, 700.0, adaptive=(0.04, 0.13, "forces", "tanh", 0.1)),
        "adaptive_force_bias_exp": lambda: case_fbmc(9003, 3, None, 500.0, adaptive=(0.02, 0.2, "energy", "exp", 0.3)),
    }
    only = set(sys.argv[1:])
    failures = []
    for name, fn in cases.items():
        if only and name not in only:
            continue
        data = fn()
        if REAL_ASE:  # compare, never write
            want = np.load(GOLDEN / f"{name}.npz", allow_pickle=True)
            identical = True
            for key in want.files:
                if key == "n_calculations" or key not in data:
                    continue
                a, b = np.asarray(data[key]), np.asarray(want[key])
                if a.shape != b.shape:
                    failures.append(f"{name}.{key}: shape {a.shape} != {b.shape}")
                elif a.dtype.kind in "iub" or b.dtype.kind in "iubSU":
                    if not np.array_equal(a, b):
                        failures.append(f"{name}.{key}: integer / string trace differs")
                elif not np.allclose(a, b, rtol=1e-10, atol=1e-12, equal_nan=True):
                    failures.append(f"{name}.{key}: max |diff| {np.nanmax(np.abs(a - b)):.3e}")
                else:
                    identical &= bool(np.array_equal(a, b, equal_nan=True))
            print(f"{name}: real ASE {'byte-identical' if identical else 'equal to 1e-10'} to the committed golden trace"
                  if not any(f.startswith(name + ".") for f in failures) else f"{name}: DIFFERS")
            continue
        np.savez_compressed(GOLDEN / f"{name}.npz", **data)
        acc = data["accepted"]
        print(f"{name}: {len(acc)} attempts, accepted {int((acc == 1).sum())}, rejected {int((acc == 0).sum())}, failed {int((acc == -1).sum())}, "
              f"E_last = {data['energy'][-1]:.12f}")  # fmt: skip
    if failures:
        print("\n".join(failures))
        raise SystemExit(1)


if __name__ == "__main__":
    main()
