"""GPU: the SM-wide balanced LEAN kernel (rollout_lean_sm_kernel: one 16-warp CTA per SM, the seventh tile of 64
rollouts cut in time between two pairs of warps) against the oracle and against the per-tile kernels.

The engine picks it on its own only when 13-14 warps' worth of rollouts land on an SM (BASELINE config 2);
MPPI_B200_BLOCK=512 forces it at any size that fits one wave, so that every model / mode / ragged shape is covered
at sizes the oracle finishes in seconds."""
import numpy as np
import pytest

from oracle import mppi_oracle as orc

pytestmark = pytest.mark.gpu
PARK = np.array([0.0, -1.0, 0.0])


def mp():
    import motion_planning_b200 as m
    return m


def make(monkeypatch, block, **kw):
    if block:
        monkeypatch.setenv("MPPI_B200_BLOCK", str(block))
    else:
        monkeypatch.delenv("MPPI_B200_BLOCK", raising=False)
    m = mp().MPPI(**kw)
    monkeypatch.delenv("MPPI_B200_BLOCK", raising=False)
    return m


@pytest.mark.parametrize("K,T", [(448 * 3, 32), (1000, 30), (64 * 6 + 1, 32), (7 * 64 * 5 + 37, 64), (40, 8 * 4 + 2)])
def test_forced_sm_wide_matches_oracle_on_ragged_shapes(monkeypatch, K, T):
    """Partially filled tiles, tiles past the last rollout, the shared tile empty / partly filled, T = 4n + 2."""
    m = make(monkeypatch, 512, horizon=T, samples=K, precision="mixed", seed=11)
    assert m.launch_info()["block"] == 512 and m.launch_info()["variant"] == "lean"
    p = orc.Params(K=K, T=T)
    s, U = np.array([0.02, -0.01, 0.1]), np.zeros((2, T))
    for _ in range(3):
        s_in = s.copy()
        s = m.get_path(s_in, PARK)
        out = orc.step(p, s_in, PARK, U, m.get_noise())
        np.testing.assert_allclose(m.latest_uvec, out["U_shift"], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(s, out["x_next"], rtol=1e-9, atol=1e-12)
        U = out["U_shift"]
    st = m.stats()
    assert st["refine_overflow"] == 0 and st["refine_max_dev"] < 5e-3
    m.close()


@pytest.mark.parametrize("split", [4, 20, 28])
def test_split_point_does_not_change_the_result(monkeypatch, split):
    """Where the shared tile is cut in time is a scheduling decision: bit-identical controls for any split."""
    K, T = 448 * 2, 32
    outs = []
    for sp in (None, split):
        if sp is None:
            monkeypatch.delenv("MPPI_B200_SPLIT", raising=False)
        else:
            monkeypatch.setenv("MPPI_B200_SPLIT", str(sp))
        m = make(monkeypatch, 512, horizon=T, samples=K, precision="f32", seed=4)
        s = np.zeros(3)
        for _ in range(2):
            s = m.get_path(s, PARK)
        outs.append((s.copy(), m.latest_uvec))
        m.close()
    monkeypatch.delenv("MPPI_B200_SPLIT", raising=False)
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("precision", ["mixed", "f32"])
def test_sm_wide_equals_per_tile_kernel(monkeypatch, precision):
    """Same tiles, same per-tile arithmetic, same partial records: block 512 and block 64 must agree bit for bit."""
    K, T = 448 * 4 + 100, 32
    res = []
    for block in (64, 512):
        m = make(monkeypatch, block, horizon=T, samples=K, precision=precision, seed=8)
        assert m.launch_info()["block"] == block
        s = np.array([0.0, 0.0, 0.3])
        for _ in range(3):
            s = m.get_path(s, PARK)
        res.append((s.copy(), m.latest_uvec))
        m.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])


def test_sm_wide_models_and_grid(monkeypatch):
    """Bicycle, unicycle-Euler and the occupancy-grid term through the SM-wide kernel, against the oracle."""
    from test_gpu_parity import demo_grid
    # bicycle
    K, T = 448 * 2 + 10, 32
    um, ns = np.array([0.22, 0.6]), np.array([0.08, 0.25])
    m = make(monkeypatch, 512, model=mp().bicycle_rk4, horizon=T, samples=K, precision="mixed", u_max=um, noise_std=ns, seed=2)
    assert m.launch_info()["block"] == 512
    p = orc.Params(K=K, T=T, model=orc.MODEL_BICYCLE, u_max=um, noise_std=ns)
    s, goal, U = np.zeros(3), np.array([1.0, 0.0, 0.0]), np.zeros((2, T))
    for _ in range(2):
        s_in = s.copy()
        s = m.get_path(s_in, goal)
        out = orc.step(p, s_in, goal, U, m.get_noise())
        np.testing.assert_allclose(m.latest_uvec, out["U_shift"], rtol=1e-8, atol=1e-9)
        U = out["U_shift"]
    m.close()
    # unicycle + Euler (theta is carried as a plain sum: start beyond pi)
    K, T = 500, 32
    m = make(monkeypatch, 512, model=mp().euler, horizon=T, samples=K, precision="mixed", u_max=[0.5, 2.0], seed=9)
    assert m.launch_info()["block"] == 512
    p = orc.Params(K=K, T=T, model=orc.MODEL_UNICYCLE_EULER, u_max=np.array([0.5, 2.0]))
    s, goal, U = np.array([0.1, 0.0, 3.3]), np.array([0.5, 0.5, 0.0]), np.zeros((2, T))
    for _ in range(2):
        s_in = s.copy()
        s = m.get_path(s_in, goal)
        out = orc.step(p, s_in, goal, U, m.get_noise())
        np.testing.assert_allclose(m.latest_uvec, out["U_shift"], rtol=1e-8, atol=1e-9)
        U = out["U_shift"]
    m.close()
    # occupancy grid
    K, T = 448 * 3, 32
    g, res, origin, w = demo_grid(), 0.06, np.array([-0.5, -0.4]), 250.0
    m = make(monkeypatch, 512, horizon=T, samples=K, precision="mixed", seed=5)
    monkeypatch.setenv("MPPI_B200_BLOCK", "512")          # set_grid re-configures the launch
    m.set_grid(g, res, origin, w)
    monkeypatch.delenv("MPPI_B200_BLOCK", raising=False)
    assert m.launch_info()["block"] == 512
    p = orc.Params(K=K, T=T, grid=g, grid_res=res, grid_origin=origin, w_obs=w)
    s, goal, U = np.array([1.013, 1.517, 0.0]), np.array([1.8, 1.6, 0.0]), np.full((2, T), 5.0)
    m.latest_uvec = U
    for _ in range(2):
        s_in = s.copy()
        s = m.get_path(s_in, goal)
        out = orc.step(p, s_in, goal, U, m.get_noise())
        np.testing.assert_allclose(m.latest_uvec, out["U_shift"], rtol=1e-8, atol=1e-9)
        U = out["U_shift"]
    m.close()


def test_engine_picks_sm_wide_only_where_it_balances(monkeypatch):
    """13-14 warps per SM (BASELINE config 2 on 148 SMs) -> SM-wide; fewer or more -> the per-tile kernels."""
    monkeypatch.delenv("MPPI_B200_BLOCK", raising=False)
    picks = {}
    for K in (65536, 32768, 131072, 128):
        m = mp().MPPI(horizon=64, samples=K, seed=0)
        picks[K] = m.launch_info()
        m.close()
    if picks[65536]["block"] == 512:       # a 148-SM part
        assert picks[65536]["grid"] == (65536 + 447) // 448
    assert picks[32768]["block"] != 512 and picks[131072]["block"] != 512 and picks[128]["block"] != 512
