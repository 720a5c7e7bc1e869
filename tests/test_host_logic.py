"""Host-side pieces of the Python mirror that need no GPU."""
import numpy as np

from motion_planning_b200.mppi import _Log


def test_log_appends_like_concatenate():
    # the reference grows path / uvec with np.concatenate every step (control/src/mppi:95-97)
    rng = np.random.default_rng(0)
    rows = rng.normal(size=(300, 3))
    log = _Log(rows[0])
    ref = np.array([rows[0]])
    for r in rows[1:]:
        log.append(r)
        ref = np.concatenate((ref, np.array([r])))
    assert log.array.shape == (300, 3)
    assert np.array_equal(log.array, ref)
    assert np.array_equal(log.array[-1], rows[-1])


def test_log_from_array_and_row_copy():
    a = np.arange(8.0).reshape(4, 2)
    log = _Log.from_array(a)
    assert np.array_equal(log.array, a)
    row = np.array([1.0, 2.0])
    log.append(row)
    row[:] = 0.0                      # the log must hold a copy, not a reference to the caller's buffer
    assert np.array_equal(log.array[-1], [1.0, 2.0])


def test_lookahead_goal_against_brute_force():
    """Controller.track_path's goal: the point `lookahead` metres of arc length past the robot's projection onto the planner's
    polyline -- checked against a dense sampling of the polyline."""
    import numpy as np
    from motion_planning_b200.controller import lookahead_goal
    rng = np.random.RandomState(0)
    for trial in range(20):
        n = rng.randint(2, 7)
        path = np.cumsum(rng.uniform(-0.2, 0.6, size=(n, 2)), axis=0)
        seg = np.diff(path, axis=0)
        seglen = np.hypot(seg[:, 0], seg[:, 1])
        cum = np.concatenate([[0], np.cumsum(seglen)])
        s = np.linspace(0, cum[-1], 20001)
        idx = np.minimum(np.searchsorted(cum, s, side="right") - 1, n - 2)
        pts = path[idx] + ((s - cum[idx]) / seglen[idx])[:, None] * seg[idx]
        for _ in range(10):
            pos = path[rng.randint(n)] + rng.normal(size=2) * 0.15
            L = rng.uniform(0.05, 0.5)
            goal, s_proj = lookahead_goal(path, pos, L)
            d = np.hypot(pts[:, 0] - pos[0], pts[:, 1] - pos[1])
            j = int(np.argmin(d))
            assert abs(d[j] - np.hypot(*(pts[np.argmin(np.abs(s - s_proj))] - pos))) < 1e-3      # projection is a nearest point
            want = pts[np.argmin(np.abs(s - min(s_proj + L, cum[-1])))]
            assert np.hypot(goal[0] - want[0], goal[1] - want[1]) < 2e-4
            k = min(np.searchsorted(cum, min(s_proj + L, cum[-1]), side="right") - 1, n - 2)
            assert abs(goal[2] - np.arctan2(seg[k, 1], seg[k, 0])) < 1e-12
    # progress is monotone: a path that returns to its start is followed in order
    loop = [[0, 0], [1, 0], [1, 1], [0, 1], [0, 0.05]]
    g1, s1 = lookahead_goal(loop, (0.02, 0.02), 0.3, s_min=0.0)
    assert s1 < 0.1 and abs(g1[1]) < 1e-12
    g2, s2 = lookahead_goal(loop, (0.02, 0.3), 0.3, s_min=3.0)
    assert s2 > 3.0 and abs(g2[0]) < 1e-12


def test_user_model_text_is_compiled_without_a_gpu():
    """mppi_check_user_model: NVRTC compiles the caller's functor text together with the embedded kernel headers for sm_100a
    (cross-compilation, no device needed); a broken text comes back with the compiler's log."""
    import motion_planning_b200 as mp
    import user_models as um
    import pytest
    mp.UserModel(um.SKID_CUDA, cost_source=um.COST_CUDA, integrator="euler", wrap_theta=False).check()
    with pytest.raises(mp.MppiError) as ei:
        mp.UserModel("template <typename R> __device__ void mppi_user_ode(const R x[3], const R u[2], R xdot[3]) { xdot[0] = nope; }").check()
    assert ei.value.status == 1 and "nope" in str(ei.value)
    # kinematic functors (kind 1) instantiate the screen family and its fp64 re-evaluation as well
    for integ in ("rk4", "euler"):
        mp.KinematicModel(um.SLIP_KIN_CUDA, integrator=integ, **um.SLIP_KIN_BOUNDS).check()
    k = mp.KinematicModel(um.DD_KIN_CUDA, **um.DD_KIN_BOUNDS)
    assert k._screenable() and k.wrap_theta and not mp.KinematicModel(um.DD_KIN_CUDA)._screenable()
    k.wrap_theta = False                          # a kinematic functor is integrated like the built-in models: rk4 wraps
    with pytest.raises(mp.MppiError) as ei:
        k.check()
    assert ei.value.status == 1
    with pytest.raises(mp.MppiError):             # an ODE functor's text is not a kinematic functor
        mp.KinematicModel(um.DD_CUDA).check()
