"""CPU: pin the NumPy restatement (oracle/mppi_oracle.py) against outputs of the live reference.

Golden files: tests/golden/ref_*.npz, made by tests/golden/make_golden.py from the unmodified
control/src/mppi.  The noise is regenerated here from the legacy NumPy stream (np.random.seed(0),
control/src/mppi:15) exactly as the reference draws it (control/src/mppi:143-146).
"""
import glob
import os

import numpy as np
import pytest
import scipy.signal

from oracle import mppi_oracle as orc
from oracle import ref_loader

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*_k*_t*.npz")))


def test_rng_kat(golden_dir):
    """SURVEY appendix B RNG known-answer test (legacy MT19937 + polar Gaussian)."""
    g = np.load(os.path.join(golden_dir, "ref_rng_kat.npz"))
    np.random.seed(0)
    assert np.array_equal(np.random.normal(size=4), g["normal4"])
    assert np.array_equal(np.random.normal(0, .9, size=(2, 5)), g["normal_2x5"])
    assert g["normal4"][0] == 1.764052345967664


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(c)[4:-4] for c in CASES])
def test_oracle_matches_reference_golden(path):
    g = np.load(path)
    K, T = int(g["K"]), int(g["T"])
    p = orc.Params(K=K, T=T)
    np.random.seed(0)
    s = g["x0"].astype(np.float64)
    U = np.zeros((2, T))
    for it in range(g["u0"].shape[0]):
        eps = orc.draw_reference_noise(p)
        if it == 0 and "eps0" in g:
            assert np.array_equal(eps, g["eps0"])
        out = orc.step(p, s, g["goal"], U, eps)
        if it == 0 and "V0" in g:
            np.testing.assert_allclose(out["V"], g["V0"], rtol=1e-13, atol=0)
        # vectorised summation order differs from the reference's .dot() chain by O(1 ulp) of V (~4e-12);
        # the softmin amplifies that by 1/lam, hence 1e-8 rather than 1e-15.
        np.testing.assert_allclose(out["u0"], g["u0"][it], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(out["x_next"], g["x_next"][it], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(out["U_shift"], g["U_shift"][it], rtol=1e-8, atol=1e-9)
        s, U = out["x_next"], out["U_shift"]


@pytest.mark.parametrize("T", [6, 16, 32, 64, 100, 128])
def test_savgol_matrix_matches_scipy(T):
    """oracle.savgol_matrix restates scipy.signal.savgol_filter(U, T-1, 3, axis=1) (control/src/mppi:202)."""
    rng = np.random.RandomState(T)
    U = rng.normal(size=(2, T)) * 3
    want = scipy.signal.savgol_filter(U, T - 1, 3, axis=1)
    got = U @ orc.savgol_matrix(T).T
    np.testing.assert_allclose(got, want, rtol=0, atol=5e-13)


def test_wrap_range():
    th = np.array([-10.0, -np.pi, -np.pi + 1e-12, 0.0, np.pi, np.pi + 1e-12, 7.0, 100.0])
    w = orc._wrap(th)
    assert np.all(w > -np.pi - 1e-15) and np.all(w <= np.pi + 1e-15)
    np.testing.assert_allclose(np.sin(w), np.sin(th), atol=1e-12)


@pytest.mark.skipif(ref_loader.available() is None, reason="reference not loadable here")
def test_oracle_matches_live_reference_with_replayed_noise():
    """Direction (ii) of SURVEY 8c: feed OUR noise tensor into the untouched reference by patching
    np.random.normal, and compare a full get_path."""
    ref = ref_loader.load_reference()
    K, T = 64, 16
    rng = np.random.RandomState(123)
    eps = (rng.standard_normal((T, 2, K)).astype(np.float32) * np.float32(0.9)).astype(np.float64)
    m = ref.MPPI(horizon=T, samples=K)
    m.latest_uvec = rng.normal(size=(2, T))
    U0 = m.latest_uvec.copy()
    feed = iter(eps)
    orig = np.random.normal
    np.random.normal = lambda *a, **k: next(feed).copy()
    try:
        x1 = m.get_path(np.array([0.1, 0.2, 3.0]), np.array([1.0, -1.0, 0.5]))
    finally:
        np.random.normal = orig
    p = orc.Params(K=K, T=T)
    out = orc.step(p, np.array([0.1, 0.2, 3.0]), np.array([1.0, -1.0, 0.5]), U0, eps)
    np.testing.assert_allclose(out["u0"], m.uvec[-1], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(out["x_next"], x1, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(out["U_shift"], m.latest_uvec, rtol=1e-8, atol=1e-9)


@pytest.mark.skipif(ref_loader.available() is None, reason="reference not loadable here")
@pytest.mark.parametrize("seed", range(6))
def test_oracle_matches_live_reference_on_varied_parameters(seed):
    """The reference's tunables are public attributes / arguments (Q, R, P1 at control/src/mppi:69-73; sig, lam at :85):
    random diagonal Q / P1, a full R, non-default sig (whose [0,0] entry is ALSO the noise std, :144-146) and lam, random
    start / goal / nominal -- two consecutive get_path calls of the live reference against the restatement."""
    rng = np.random.RandomState(1000 + seed)
    K, T = int(rng.choice([7, 32, 65])), int(rng.choice([6, 16, 30]))
    Q = np.array([rng.uniform(1, 2e3), rng.uniform(1, 2e3), rng.choice([0.0, rng.uniform(0, 50)])])
    P1 = rng.uniform(1, 2e3, size=3)
    R = np.array([[rng.uniform(0.5, 2), 0.1], [0.1, rng.uniform(0.5, 2)]])
    sig = np.array([[rng.uniform(0.3, 1.2), 0.05], [0.02, rng.uniform(0.3, 1.2)]])
    lam = float(rng.choice([1e-3, 1e-2, 0.1]))
    x0, goal = rng.uniform(-1, 1, size=3) * [1, 1, 3], rng.uniform(-1, 1, size=3) * [1, 1, 3]
    ref = ref_loader.load_reference()
    m = ref.MPPI(horizon=T, samples=K)
    m.Q, m.R, m.P1 = np.diag(Q), R.copy(), np.diag(P1)
    m.latest_uvec = rng.normal(size=(2, T)) * 2
    p = orc.Params(K=K, T=T, Q=Q, R=R, P1=P1, sig=sig, lam=lam, noise_std=np.array([sig[0, 0], sig[0, 0]]))
    U, s = m.latest_uvec.copy(), x0.copy()
    sref = x0.copy()
    orig = np.random.normal
    for _ in range(2):
        eps = rng.standard_normal((T, 2, K)) * sig[0, 0]
        feed = iter(eps)
        np.random.normal = lambda *a, **k: next(feed).copy()
        try:
            sref = m.get_path(sref, goal, sig=sig, lam=lam)
        finally:
            np.random.normal = orig
        out = orc.step(p, s, goal, U, eps)
        np.testing.assert_allclose(out["u0"], m.uvec[-1], rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(out["x_next"], sref, rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(out["U_shift"], m.latest_uvec, rtol=1e-7, atol=1e-8)
        # continue from the reference's own state so that the second step is compared from identical inputs
        s, U = sref.copy(), m.latest_uvec.copy()


@pytest.mark.skipif(ref_loader.available() is None, reason="reference not loadable here")
@pytest.mark.parametrize("seed", range(3))
def test_oracle_euler_model_matches_live_reference(seed):
    """The reference's second integrator-step functor, `euler` over `unicycle_dynamics` (control/src/mppi:33-36,57-58),
    plugged into the UNMODIFIED class through its own `model=` hook (:62,66,154): pins the oracle's
    MODEL_UNICYCLE_EULER path (no theta wrap, explicit Euler) to the reference itself."""
    ref = ref_loader.load_reference()
    rng = np.random.RandomState(300 + seed)
    K, T = int(rng.choice([16, 48])), int(rng.choice([8, 16, 30]))
    m = ref.MPPI(model=ref.euler, horizon=T, samples=K)
    m.latest_uvec = rng.normal(size=(2, T))
    p = orc.Params(K=K, T=T, model=orc.MODEL_UNICYCLE_EULER)
    x0, goal = rng.uniform(-1, 1, size=3) * [1, 1, 3], rng.uniform(-1, 1, size=3) * [1, 1, 3]
    U, s = m.latest_uvec.copy(), x0.copy()
    orig = np.random.normal
    for _ in range(3):
        eps = rng.standard_normal((T, 2, K)) * 0.9
        feed = iter(eps)
        np.random.normal = lambda *a, **k: next(feed).copy()
        try:
            sref = m.get_path(s, goal)
        finally:
            np.random.normal = orig
        out = orc.step(p, s, goal, U, eps)
        np.testing.assert_allclose(out["u0"], m.uvec[-1], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(out["x_next"], sref, rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(out["U_shift"], m.latest_uvec, rtol=1e-8, atol=1e-9)
        s, U = sref.copy(), m.latest_uvec.copy()


@pytest.mark.skipif(ref_loader.available() is None, reason="reference not loadable here")
@pytest.mark.parametrize("integrator,wrap", [("rk4", True), ("euler", False)])
@pytest.mark.parametrize("ode_name", ["skid_numpy", "slip_numpy"])
def test_oracle_user_model_matches_live_reference_through_its_model_hook(integrator, wrap, ode_name):
    """an arbitrary ODE plugged into the UNMODIFIED reference class through its own `model=` constructor
    argument (control/src/mppi:62,66,154,213) -- here one whose speed and yaw rate depend on the state -- and, for the cost
    functor, a subclass overriding get_cost (:180-184).  Pins the oracle's MODEL_USER path and its cost hooks.  slip_numpy is
    the NumPy twin of the KINEMATIC functor the GPU tests run (speed and yaw rate from the controls only)."""
    import user_models as um
    ref = ref_loader.load_reference()
    rng = np.random.RandomState(77)
    K, T = 24, 12
    ode = getattr(um, ode_name)
    model = orc.user_model_step(ode, integrator, wrap)

    class CostMPPI(ref.MPPI):
        def get_cost(self, state, goal, u, lam, sig, eps):
            return float(um.running_cost_numpy(state.reshape(3, 1), goal, u, eps.reshape(2, 1), 0)[0])

    for cls, cost in ((ref.MPPI, False), (CostMPPI, True)):
        m = cls(model=model, horizon=T, samples=K)
        m.latest_uvec = rng.normal(size=(2, T)) * 2
        p = orc.Params(K=K, T=T, model=orc.MODEL_USER, user_ode=ode, user_integrator=integrator, user_wrap=wrap)
        if cost:
            p.user_running_cost, p.user_terminal_cost = um.running_cost_numpy, um.terminal_cost_numpy
        x0, goal = np.array([0.2, -0.1, 0.7]), np.array([0.6, -0.5, -0.3])
        U, s = m.latest_uvec.copy(), x0.copy()
        orig = np.random.normal
        for _ in range(2):
            eps = rng.standard_normal((T, 2, K)) * 0.9
            feed = iter(eps)
            np.random.normal = lambda *a, **k: next(feed).copy()
            try:
                sref = m.get_path(s, goal)
            finally:
                np.random.normal = orig
            out = orc.step(p, s, goal, U, eps)
            np.testing.assert_allclose(out["u0"], m.uvec[-1], rtol=1e-8, atol=1e-10)
            np.testing.assert_allclose(out["x_next"], sref, rtol=1e-8, atol=1e-12)
            np.testing.assert_allclose(out["U_shift"], m.latest_uvec, rtol=1e-8, atol=1e-9)
            s, U = sref.copy(), m.latest_uvec.copy()
