"""GPU: the engine against the oracle over
the reference's tunables -- diagonal Q / P1, a full R, a non-diagonal sig (whose [0,0] entry is also the noise std,
control/src/mppi:144-146) and lam -- the CUDA twin of
tests/test_oracle_golden.py::test_oracle_matches_live_reference_on_varied_parameters (which pins the oracle itself)."""
import numpy as np
import pytest

from oracle import mppi_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["f64", "mixed"])
@pytest.mark.parametrize("seed", range(6))
def test_engine_matches_oracle_on_varied_parameters(seed, precision):
    import motion_planning_b200 as mp
    rng = np.random.RandomState(2000 + seed)
    K, T = int(rng.choice([7, 320, 1000])), int(rng.choice([6, 16, 30]))
    Q = np.array([rng.uniform(1, 2e3), rng.uniform(1, 2e3), rng.choice([0.0, rng.uniform(0, 50)])])
    if seed % 2 == 0:
        Q[1], Q[2] = Q[0], 0.0                      # admissible for the LEAN path (given the yaw-increment bound)
    P1 = rng.uniform(1, 2e3, size=3)
    R = np.array([[rng.uniform(0.5, 2), 0.1], [0.1, rng.uniform(0.5, 2)]])
    sig = np.array([[rng.uniform(0.3, 1.2), 0.05], [0.02, rng.uniform(0.3, 1.2)]])
    lam = float(rng.choice([1e-3, 1e-2, 0.1]))
    x0, goal = rng.uniform(-1, 1, size=3) * [1, 1, 3], rng.uniform(-1, 1, size=3) * [1, 1, 3]
    m = mp.MPPI(horizon=T, samples=K, precision=precision, seed=seed)
    m.Q, m.R, m.P1 = np.diag(Q), R.copy(), np.diag(P1)
    U = rng.normal(size=(2, T)) * 2
    m.get_path(x0, goal, sig=sig, lam=lam)          # applies Q / R / P1 / sig / lam (re-creates the engine)
    m.latest_uvec = U
    p = orc.Params(K=K, T=T, Q=Q, R=R, P1=P1, sig=sig, lam=lam, noise_std=np.array([sig[0, 0], sig[0, 0]]))
    s = x0.copy()
    for _ in range(2):
        s_in = s.copy()
        s = m.get_path(s_in, goal, sig=sig, lam=lam)
        out = orc.step(p, s_in, goal, U, m.get_noise())
        np.testing.assert_allclose(m.latest_uvec, out["U_shift"], rtol=1e-7, atol=1e-8)
        np.testing.assert_allclose(s, out["x_next"], rtol=1e-8, atol=1e-11)
        U = out["U_shift"]
    m.close()
