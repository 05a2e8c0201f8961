"""The CUDA path against the golden traces of the GENUINE quansino sources (tests/golden/, oracle/gen_golden.py)."""

import numpy as np
import pytest

from tests.golden_cases import CASES, gpu_replay, load

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_variant")]


@pytest.mark.parametrize("name", CASES)
def test_cuda_reproduces_genuine_quansino(cuda_lib, name):
    g = load(name)
    mc = gpu_replay(g)
    tr = mc.last_trace
    assert np.array_equal(tr["accepted"][0], g["accepted"]), "accept / reject / failed sequence differs from quansino"
    assert np.array_equal(tr["move"][0], g["move"])
    np.testing.assert_allclose(tr["energy"][0], g["energy"], rtol=1e-10, atol=1e-10)
    out = mc.download()
    n = int(out.n_atoms[0])
    assert n == g["n_atoms"][-1]
    np.testing.assert_allclose(out.positions[0, :n], g["positions"], rtol=0, atol=1e-10)
    if g["kind"] in ("isobaric", "hybrid", "forced"):
        np.testing.assert_allclose(out.cell[0], g["cell"], rtol=1e-13)
    if g["kind"] in ("gc_lj", "gc_emt"):
        assert mc.number_of_exchange_particles[0] == g["n_exchange"]
        assert np.array_equal(mc.labels("default_displacement_move")[0, :n], g["labels_displacement"])
        assert np.array_equal(mc.labels("default_exchange_move")[0, :n], g["labels_exchange"])
    if g["kind"] == "hmc":
        np.testing.assert_allclose(out.momenta[0], g["momenta"], rtol=0, atol=1e-9)
