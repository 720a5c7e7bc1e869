"""CPU: pin the C/OpenMP port (oracle/mppi_port.c) against the golden vectors of the live reference."""
import glob
import os

import numpy as np
import pytest

from oracle import build_port, port_c

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*_k*_t*.npz")))


@pytest.fixture(scope="module", autouse=True)
def built():
    build_port.build()


def test_legacy_mt19937_gaussian_is_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_rng_kat.npz"))
    port = port_c.Port(4, 6, seed=0)
    assert np.array_equal(port.normal(1.0, 4), g["normal4"])
    assert np.array_equal(port.normal(0.9, 10).reshape(2, 5), g["normal_2x5"])


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(c)[4:-4] for c in CASES])
def test_port_reproduces_reference_closed_loop(path):
    """Self-generated legacy noise stream (seed 0) + the reference's op order: the whole closed loop."""
    g = np.load(path)
    K, T = int(g["K"]), int(g["T"])
    port = port_c.Port(K, T, seed=0)
    s = g["x0"].astype(np.float64)
    for it in range(g["u0"].shape[0]):
        out = port.step(s, g["goal"], noise_mode=1)
        if it == 0 and "eps0" in g:
            assert np.array_equal(out["eps"], g["eps0"])
        np.testing.assert_allclose(out["u0"], g["u0"][it], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(out["x_next"], g["x_next"][it], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(out["U_shift"], g["U_shift"][it], rtol=1e-8, atol=1e-9)
        s = out["x_next"]


def test_port_value_function(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_c1_park_k128_t32.npz"))
    port = port_c.Port(128, 32)
    V = port.cost2go(g["x0"], np.zeros((2, 32)), g["goal"], g["eps0"])
    np.testing.assert_allclose(V, g["V0"], rtol=1e-13)
