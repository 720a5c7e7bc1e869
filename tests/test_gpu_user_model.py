"""GPU: caller-supplied dynamics / cost functors (mppi_create_user): CUDA text compiled at run time into its
own instantiation of the rollout / reduce / finalize kernels.  Checked against the oracle's MODEL_USER path -- which
tests/test_oracle_golden.py pins to the UNMODIFIED reference class through its own `model=` hook -- and, where
oracle/_ref/mppi.pyc travelled, against that reference class itself with the GPU's noise replayed into it."""
import numpy as np
import pytest

import user_models as um
from oracle import mppi_oracle as orc
from oracle import ref_loader

pytestmark = pytest.mark.gpu
PARK = np.array([0.0, -1.0, 0.0])


def mp():
    import motion_planning_b200 as m
    return m


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def test_user_diff_drive_functor_reproduces_the_builtin_model():
    """dd_dynamics (control/src/mppi:23-30) written as a user ODE and run through the generic RK4 must give what the built-in
    diff-drive kernels give (which take the yaw-rate-independent short cut): V to 1e-11, the control sequence to 1e-8."""
    K, T = 2048, 32
    user = mp().MPPI(model=mp().UserModel(um.DD_CUDA, name="dd_user"), horizon=T, samples=K, precision="f64", seed=4)
    built = mp().MPPI(horizon=T, samples=K, precision="f64", seed=4)
    p = orc.Params(K=K, T=T)
    for m in (user, built):
        m.set_capture(True)
    s, U = np.array([0.1, 0.0, 0.4]), np.zeros((2, T))
    for it in range(3):
        s_in = s.copy()
        su = user.get_path(s_in, PARK)
        s = built.get_path(s_in, PARK)
        np.testing.assert_allclose(user.get_value_fcn(), built.get_value_fcn(), rtol=1e-11)
        np.testing.assert_allclose(user.latest_uvec, built.latest_uvec, rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(su, s, rtol=1e-10, atol=1e-13)
        out = orc.step(p, s_in, PARK, U, user.get_noise())
        np.testing.assert_allclose(user.latest_uvec, out["U_shift"], rtol=1e-8, atol=1e-9)
        U = out["U_shift"]
        user.latest_uvec = built.latest_uvec
    user.close()
    built.close()


@pytest.mark.parametrize("integrator,wrap", [("rk4", True), ("euler", False)])
@pytest.mark.parametrize("with_cost", [False, True])
def test_user_model_against_oracle_and_reference(integrator, wrap, with_cost):
    """a model none of the built-in kernels can express (state-dependent speed and yaw rate), optionally with a cost functor
    (the reference's running cost + a repulsive potential)."""
    K, T = 512, 16
    model = mp().UserModel(um.SKID_CUDA, name="skid", integrator=integrator, wrap_theta=wrap,
                           cost_source=um.COST_CUDA if with_cost else None)
    m = mp().MPPI(model=model, horizon=T, samples=K, seed=9)          # precision defaults to f64 for a user model
    m.set_capture(True)
    p = orc.Params(K=K, T=T, model=orc.MODEL_USER, user_ode=um.skid_numpy, user_integrator=integrator, user_wrap=wrap)
    if with_cost:
        p.user_running_cost, p.user_terminal_cost = um.running_cost_numpy, um.terminal_cost_numpy
    ref = None
    if ref_loader.available():
        R = ref_loader.load_reference()
        cls = R.MPPI
        if with_cost:
            class CostMPPI(R.MPPI):
                def get_cost(self, state, goal, u, lam, sig, eps):
                    return float(um.running_cost_numpy(state.reshape(3, 1), goal, u, eps.reshape(2, 1), 0)[0])
            cls = CostMPPI
        ref = cls(model=orc.user_model_step(um.skid_numpy, integrator, wrap), horizon=T, samples=K)
    s, goal, U = np.array([0.2, -0.1, 0.7]), np.array([0.6, -0.5, -0.3]), np.zeros((2, T))
    for it in range(3):
        s_in = s.copy()
        s = m.get_path(s_in, goal)
        eps = m.get_noise()
        out = orc.step(p, s_in, goal, U, eps)
        np.testing.assert_allclose(m.get_value_fcn(), out["V"], rtol=1e-11)
        np.testing.assert_allclose(m.uvec[-1], out["u0"], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(s, out["x_next"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(m.latest_uvec, out["U_shift"], rtol=1e-8, atol=1e-9)
        if ref is not None:
            ref.latest_uvec = U.copy()
            feed, orig = iter(eps), np.random.normal
            np.random.normal = lambda *a, **k: next(feed).copy()
            try:
                xr = ref.get_path(s_in, goal)
            finally:
                np.random.normal = orig
            np.testing.assert_allclose(m.uvec[-1], ref.uvec[-1], rtol=1e-8, atol=1e-9)
            np.testing.assert_allclose(s, xr, rtol=1e-8, atol=1e-11)
            np.testing.assert_allclose(m.latest_uvec, ref.latest_uvec, rtol=1e-8, atol=1e-8)
        U = out["U_shift"]
    # the `model` functor itself (control/src/mppi:154): model(states, u, dt) on the device
    rng = np.random.RandomState(0)
    xs, us = rng.normal(size=(3, 40)) * 2, rng.normal(size=(2, 40)) * 3
    np.testing.assert_allclose(model(xs, us, 1.0 / T), orc.user_model_step(um.skid_numpy, integrator, wrap)(xs, us, 1.0 / T), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(m.perform_action(s, m.latest_uvec), orc.perform_action(p, s, m.latest_uvec), rtol=1e-12, atol=1e-14)
    m.close()


def test_user_model_fp32_and_grid_and_errors():
    M = mp()
    K, T = 4096, 32
    model = M.UserModel(um.SKID_CUDA, name="skid")
    a = M.MPPI(model=model, horizon=T, samples=K, precision="f64", seed=2)
    b = M.MPPI(model=model, horizon=T, samples=K, precision="f32", seed=2)
    g = np.zeros((40, 40), dtype=np.int8)
    g[18:24, 24:30] = 100
    for m in (a, b):
        m.set_grid(g, 0.05, np.array([-1.0, -1.0]), 50.0)
    p = orc.Params(K=K, T=T, model=orc.MODEL_USER, user_ode=um.skid_numpy, grid=g, grid_res=0.05, grid_origin=np.array([-1.0, -1.0]), w_obs=50.0)
    s_in, goal = np.array([0.0, 0.0, 0.2]), np.array([0.8, 0.1, 0.0])
    U0 = np.full((2, T), 5.0)
    a.latest_uvec = U0
    b.latest_uvec = U0
    a.get_path(s_in, goal)
    b.get_path(s_in, goal)
    out = orc.step(p, s_in, goal, U0, a.get_noise())
    np.testing.assert_allclose(a.latest_uvec, out["U_shift"], rtol=1e-8, atol=1e-9)
    assert np.array_equal(a.get_noise(), b.get_noise())
    assert rel_err(b.latest_uvec, a.latest_uvec) < 5e-2           # the literal fp32 pipeline: conditioning-limited
    with pytest.raises(M.MppiError) as ei:
        M.MPPI(model=model, horizon=T, samples=K, precision="mixed")
    assert ei.value.status == 4                                   # MPPI_ERR_UNSUPPORTED
    with pytest.raises(M.MppiError) as ei:
        M.MPPI(model=M.UserModel("template <typename R> __device__ void mppi_user_ode(const R x[3], const R u[2], R xdot[3]) { xdot[0] = nope; }"),
               horizon=T, samples=K)
    assert ei.value.status == 1 and "nope" in str(ei.value)
    a.close()
    b.close()
