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


def test_kinematic_functor_restating_the_builtin_diff_drive_reproduces_it(monkeypatch):
    """mppi_user_model.kind 1: the functor replaces speed_yaw() inside the BUILT-IN kernels.  dd_dynamics (control/src/mppi:23-30)
    as a kinematic functor, against the built-in diff-drive pinned to the same (general) code path: same instructions around
    the functor, so precision 'mixed' must agree to rounding of the functor's own constants -- far inside 1e-9 -- and both
    self-checks of the screen must hold."""
    M = mp()
    K, T = 8192, 32
    kin = M.KinematicModel(um.DD_KIN_CUDA, name="dd_kin", **um.DD_KIN_BOUNDS)
    user = M.MPPI(model=kin, horizon=T, samples=K, seed=4)         # precision defaults to 'mixed' for a screenable functor
    user64 = M.MPPI(model=kin, horizon=T, samples=K, precision="f64", seed=4)
    monkeypatch.setenv("MPPI_B200_VARIANT", "general")
    built = M.MPPI(horizon=T, samples=K, precision="mixed", seed=4)
    monkeypatch.delenv("MPPI_B200_VARIANT")
    user64.set_capture(True)
    p = orc.Params(K=K, T=T)
    s, U = np.array([0.1, 0.0, 0.4]), np.zeros((2, T))
    for it in range(4):
        s_in = s.copy()
        su = user.get_path(s_in, PARK)
        s64 = user64.get_path(s_in, PARK)
        s = built.get_path(s_in, PARK)
        assert rel_err(user.latest_uvec, built.latest_uvec) < 1e-9
        assert rel_err(user.latest_uvec, user64.latest_uvec) < 1e-9
        np.testing.assert_allclose(su, s, rtol=0, atol=1e-12)
        np.testing.assert_allclose(s64, s, rtol=0, atol=1e-12)
        st, sb = user.stats(), built.stats()
        assert st["refine_overflow"] == 0 and st["refine_candidates"] >= T
        assert st["refine_candidates"] == sb["refine_candidates"]             # the same screen on the same noise
        assert 0 < st["refine_head_room"] < 0.1 and st["refine_max_dev"] < st["refine_head_room"] / 4, st
        assert abs(st["refine_head_room"] - sb["refine_head_room"]) < 1e-9 * sb["refine_head_room"]   # the stated bounds == the built-in's
        out = orc.step(p, s_in, PARK, U, user64.get_noise())
        np.testing.assert_allclose(user.latest_uvec, out["U_shift"], rtol=1e-8, atol=1e-9)
        U = out["U_shift"]
        for m in (user, user64):
            m.latest_uvec = built.latest_uvec
    for m in (user, user64, built):
        m.close()


@pytest.mark.parametrize("integrator", ["rk4", "euler"])
def test_kinematic_functor_of_a_new_vehicle_all_precisions(integrator):
    """a vehicle no built-in model is (slipping tracks), at a rollout count where the screen matters: f64 against the oracle's
    generic integrator (the model as an ODE) and against the reference class where it travelled; mixed == f64 at 1e-9 with the
    screen's self-checks; f32 within the conditioning bound; a grid in play."""
    M = mp()
    K, T = 16384, 32
    kin = M.KinematicModel(um.SLIP_KIN_CUDA, name="slip", integrator=integrator, **um.SLIP_KIN_BOUNDS)
    wrap = integrator == "rk4"
    g = np.zeros((40, 40), dtype=np.int8)
    g[18:24, 24:30] = 100
    grid = dict(grid=g, grid_res=0.05, grid_origin=np.array([-1.0, -1.0]), w_obs=50.0)
    eng = {}
    for prec in ("f64", "mixed", "f32"):
        eng[prec] = M.MPPI(model=kin, horizon=T, samples=K, precision=prec, seed=6)
        eng[prec].set_grid(g, 0.05, np.array([-1.0, -1.0]), 50.0)
    eng["f64"].set_capture(True)
    p = orc.Params(K=K, T=T, model=orc.MODEL_USER, user_ode=um.slip_numpy, user_integrator=integrator, user_wrap=wrap, **grid)
    s, goal = np.array([0.0, 0.0, 0.2]), np.array([0.8, 0.1, 0.0])
    U = np.full((2, T), 5.0)
    for m in eng.values():
        m.latest_uvec = U
    for it in range(3):
        s_in = s.copy()
        s = eng["f64"].get_path(s_in, goal)
        sm = eng["mixed"].get_path(s_in, goal)
        eng["f32"].get_path(s_in, goal)
        out = orc.step(p, s_in, goal, U, eng["f64"].get_noise())
        np.testing.assert_allclose(eng["f64"].get_value_fcn(), out["V"], rtol=1e-10)
        np.testing.assert_allclose(eng["f64"].latest_uvec, out["U_shift"], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(s, out["x_next"], rtol=1e-9, atol=1e-12)
        assert rel_err(eng["mixed"].latest_uvec, eng["f64"].latest_uvec) < 1e-9
        np.testing.assert_allclose(sm, s, rtol=0, atol=1e-12)
        st = eng["mixed"].stats()
        assert st["refine_overflow"] == 0 and st["refine_candidates"] >= T
        assert 0 < st["refine_head_room"] < 0.1 and st["refine_max_dev"] < st["refine_head_room"] / 4, st
        assert rel_err(eng["f32"].latest_uvec, eng["f64"].latest_uvec) < 5e-2
        U = out["U_shift"]
        eng["f32"].latest_uvec = eng["f64"].latest_uvec
        eng["mixed"].latest_uvec = eng["f64"].latest_uvec
    if ref_loader.available() and integrator == "rk4":
        R = ref_loader.load_reference()
        ref = R.MPPI(model=orc.user_model_step(um.slip_numpy, integrator, wrap), horizon=T, samples=512)
        small = M.MPPI(model=kin, horizon=T, samples=512, seed=1)              # mixed
        small.set_capture(True)
        s0, g0 = np.array([0.2, -0.1, 0.7]), np.array([0.6, -0.5, -0.3])
        xg = small.get_path(s0, g0)
        feed, orig = iter(small.get_noise()), np.random.normal
        np.random.normal = lambda *a, **k: next(feed).copy()
        try:
            xr = ref.get_path(s0, g0)
        finally:
            np.random.normal = orig
        np.testing.assert_allclose(xg, xr, rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(small.latest_uvec, ref.latest_uvec, rtol=1e-8, atol=1e-8)
        small.close()
    # perform_action / the model functor with the kinematic functor
    rng = np.random.RandomState(0)
    xs, us = rng.normal(size=(3, 40)) * 2, rng.normal(size=(2, 40)) * 3
    np.testing.assert_allclose(kin(xs, us, 1.0 / T), orc.user_model_step(um.slip_numpy, integrator, wrap)(xs, us, 1.0 / T), rtol=1e-11, atol=1e-13)
    for m in eng.values():
        m.close()


def test_kinematic_functor_errors_and_understated_bounds():
    M = mp()
    K, T = 4096, 32
    with pytest.raises(M.MppiError) as ei:       # mixed without bounds
        M.MPPI(model=M.KinematicModel(um.SLIP_KIN_CUDA), horizon=T, samples=K, precision="mixed")
    assert ei.value.status == 1
    with pytest.raises(M.MppiError) as ei:       # mixed with a cost functor: the screen's delta-form cost is the built-in one
        M.MPPI(model=M.KinematicModel(um.SLIP_KIN_CUDA, cost_source=um.COST_CUDA, **um.SLIP_KIN_BOUNDS), horizon=T, samples=K, precision="mixed")
    assert ei.value.status == 4
    # bounds understated 100x: the window's head-room shrinks; whatever the screen then does, the safety net (a step whose
    # measured fp32 error exceeds half the head-room is redone in fp64) keeps the result equal to f64
    lying = M.KinematicModel(um.SLIP_KIN_CUDA, speed_max=um.SLIP_KIN_BOUNDS["speed_max"] / 100, yaw_rate_max=um.SLIP_KIN_BOUNDS["yaw_rate_max"] / 100)
    a = M.MPPI(model=lying, horizon=T, samples=K, precision="f64", seed=3)
    b = M.MPPI(model=lying, horizon=T, samples=K, precision="mixed", seed=3)
    s = np.zeros(3)
    for it in range(3):
        sa = a.get_path(s, PARK)
        sb = b.get_path(s, PARK)
        assert rel_err(b.latest_uvec, a.latest_uvec) < 1e-9
        np.testing.assert_allclose(sb, sa, rtol=0, atol=1e-12)
        b.latest_uvec = a.latest_uvec
        s = sa
    a.close()
    b.close()
