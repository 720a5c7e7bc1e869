"""GPU: the three code paths of the rollout kernel (general / fast / lean) agree, and the engine picks the
one its parameters admit (engine.cu try_configure; rollout_lean_kernel.cuh header lists the conditions)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
PARK = np.array([0.0, -1.0, 0.0])


def mp():
    import motion_planning_b200 as m
    return m


def _run(monkeypatch, variant, precision, K=4096, T=32, steps=3, **kw):
    if variant:
        monkeypatch.setenv("MPPI_B200_VARIANT", variant)
    else:
        monkeypatch.delenv("MPPI_B200_VARIANT", raising=False)
    m = mp().MPPI(horizon=T, samples=K, seed=3, precision=precision, **kw)
    info = m.launch_info()
    s = np.array([0.05, 0.0, 0.2])
    out = []
    for _ in range(steps):
        s = m.get_path(s, PARK)
        out.append((m.uvec[-1].copy(), s.copy(), m.latest_uvec, m.stats()))
    m.close()
    return info, out


def test_mixed_result_does_not_depend_on_the_rollout_code_path(monkeypatch):
    """mixed precision: the fp32 rollouts only SCREEN; the soft-min is formed in fp64 from the same Philox counters,
    so general / fast / lean must produce the same controls to fp64 rounding."""
    ref_info, ref = _run(monkeypatch, "general", "mixed")
    assert ref_info["variant"] == "general"
    for variant in ("fast", None):
        info, out = _run(monkeypatch, variant, "mixed")
        assert info["variant"] == (variant or "lean")
        for (u, s, U, st), (u0, s0, U0, st0) in zip(out, ref):
            np.testing.assert_allclose(U, U0, rtol=1e-9, atol=1e-11)
            np.testing.assert_allclose(u, u0, rtol=1e-9, atol=1e-11)
            np.testing.assert_allclose(s, s0, rtol=1e-9, atol=1e-12)
            assert st["refine_overflow"] == 0 and st["refine_max_dev"] < 5e-3   # fp32 screen within its head-room


def test_f32_code_paths_agree_within_fp32_rounding(monkeypatch):
    """precision f32 (fp32 online soft-min): the paths differ only by fp32 rounding of the cost-to-go; compare the
    first step (same nominal, same noise) with the tolerance of an fp32 soft-min at lambda = 1e-3."""
    _, ref = _run(monkeypatch, "general", "f32", steps=1)
    for variant in ("fast", None):
        _, out = _run(monkeypatch, variant, "f32", steps=1)
        np.testing.assert_allclose(out[0][2], ref[0][2], rtol=0, atol=2e-2)   # a few fp32 ulps of V/lambda on |U| ~ 1


@pytest.mark.parametrize("kw,T,expect", [
    ({}, 32, "lean"),                                   # reference defaults: Q[2] = 0, dt*yaw <= 1/8
    ({}, 8, "fast"),                                    # dt = 1/8: |dt*yaw| = 0.33 > 1/8 but <= pi/4
    ({"u_max": 20.0}, 8, "general"),                    # wheel speeds up to 20 rad/s: |dt*yaw| = 1.03 > pi/4
])
def test_engine_picks_the_admissible_path(monkeypatch, kw, T, expect):
    monkeypatch.delenv("MPPI_B200_VARIANT", raising=False)
    m = mp().MPPI(horizon=T, samples=512, seed=0, **kw)
    assert m.launch_info()["variant"] == expect
    m.get_path(np.zeros(3), PARK)
    m.close()


def test_theta_cost_falls_back_to_the_fast_path(monkeypatch):
    """Q[2] != 0: theta enters the running cost, the lean kernel (which wraps theta lazily) is not admissible."""
    monkeypatch.delenv("MPPI_B200_VARIANT", raising=False)
    m = mp().MPPI(horizon=32, samples=512, seed=0)
    m.Q = np.diag([1e3, 1e3, 10.0])
    m.get_path(np.zeros(3), PARK)          # re-creates the engine with the new cost
    assert m.launch_info()["variant"] == "fast"
    m.close()
