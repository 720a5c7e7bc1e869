"""CPU: the ROS-free `Controller` mirror (motion_planning_b200/controller.py) against the UNMODIFIED reference
node (control/src/mppi:296-389, driven by oracle/ref_controller.py) and against the committed golden traces.

The state machine is host logic, so it is exercised here with the reference's own MPPI object plugged into
the mirror: any difference in the twists is then a difference in the Controller, not in the hot path."""
import os

import numpy as np
import pytest

from motion_planning_b200.controller import Controller, FakeDiffDrive, run_closed_loop, waypoints_from_path, yaw_from_quaternion
from oracle import ref_controller, ref_loader

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(ref_loader.available() is None, reason="reference MPPI not loadable (no source, no oracle/_ref/mppi.pyc)")

CASES = [("waypoints_k32_t16", (0.0, 0.0, 0.0)), ("park_k32_t16", (0.0, -0.93, 0.4))]


def _golden(name):
    g = np.load(os.path.join(GOLDEN, "node_%s.npz" % name))
    wps = g["waypoints"].tolist() if g["waypoints"].size else None
    return g, wps, int(g["K"]), int(g["T"])


@needs_ref
@pytest.mark.parametrize("name,pose0", CASES)
def test_mirror_with_reference_mppi_reproduces_golden_trace(name, pose0):
    """Same MPPI object class, same noise stream => the mirror must publish the reference node's twists bit for bit
    (odometry delivered as Odometry-shaped messages, like the node receives them)."""
    g, wps, K, T = _golden(name)
    ref = ref_loader.load_reference()                  # re-seeds the global stream (control/src/mppi:15)
    sent = []
    node = Controller(mppi=ref.MPPI(horizon=T, samples=K), waypoints=wps, publish=lambda vx, wz: sent.append((vx, wz)))
    assert sent == [(0.0, 0.0)]                         # the start-up Twist (control/src/mppi:313-316)
    plant = FakeDiffDrive(pose0, dt=1.0 / T)
    flags = []
    for i in range(len(g["poses"])):
        # the recorded odometry is replayed; the plant cross-checks it
        np.testing.assert_allclose(plant.pose, g["poses"][i], rtol=0, atol=1e-12)
        vx, wz = node.pos_cb(ref_controller.make_odom(*g["poses"][i]))
        flags.append((node.idx, int(node.init), int(node.done)))
        plant.pose = g["poses"][i].copy()
        plant.step(vx, wz)
    assert np.array_equal(np.array(sent[1:]), g["twists"])
    assert np.array_equal(np.array(flags), g["flags"])
    assert node.mppi.uvec.shape[0] == int(g["uvec_rows"])


@needs_ref
def test_mirror_matches_live_reference_node_on_the_pentagon():
    """control/config/waypoints.yaml:1 with the live reference node and the mirror side by side (plain pose triples in)."""
    wps = [[1, 0], [2, 1], [1, 2], [0, 2], [0, 0]]
    kw = dict(horizon=8, samples=6)
    _, node, log = ref_controller.load_node(wps, kw)
    plant = FakeDiffDrive((0.2, -0.1, 0.5), dt=1.0 / 8)
    poses, twists, flags, _ = ref_controller.run_node(node, plant.step, plant.pose, 60)
    ref = ref_loader.load_reference()
    mylog = []
    mine = Controller(mppi=ref.MPPI(**kw), waypoints=wps, log=mylog.append)
    poses2, twists2 = run_closed_loop(mine, FakeDiffDrive((0.2, -0.1, 0.5), dt=1.0 / 8), 60)
    np.testing.assert_allclose(twists2, twists, rtol=0, atol=1e-8)   # yaw through a quaternion vs. passed directly
    np.testing.assert_allclose(poses2, poses, rtol=0, atol=1e-8)
    assert mylog == log


def test_yaw_from_quaternion_matches_the_tf_matrix_route():
    rng = np.random.RandomState(3)
    for _ in range(200):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        assert abs(yaw_from_quaternion(*q) - ref_controller.euler_from_quaternion(q)[2]) < 1e-12
    for th in np.linspace(-np.pi + 1e-9, np.pi, 41):
        q = ref_controller.quaternion_from_yaw(th)
        assert abs(yaw_from_quaternion(*q) - th) < 1e-12


def test_wheels_to_twist_and_parallel_park_default():
    class Stub(object):
        start = np.zeros(3)
        goal = np.zeros(3)
        thresh = 0.05
        uvec = np.array([[0.0, 0.0]])

        def initialize(self):
            self.uvec = np.array([[0.0, 0.0]])

        def get_path(self, s, g):
            self.uvec = np.vstack([self.uvec, [1.0, 3.0]])
            return s

    c = Controller(mppi=Stub(), waypoints=[])
    assert c.parallel_park and c.cmd == (0.0, 0.0)
    vx, wz = c.wheelsToTwist([2.0, 4.0])
    assert vx == pytest.approx(0.033 * 3.0) and wz == pytest.approx(0.033 * 2.0 / 0.16)
    assert c.pos_cb((0.0, 0.0, 0.0)) == (0.0, 0.0)                     # first callback only initialises
    assert np.array_equal(c.mppi.goal, [0.0, -1.0, 0.0])               # control/src/mppi:337
    vx, wz = c.pos_cb((0.0, 0.0, 0.0))                                  # second one steps
    assert (vx, wz) == c.wheelsToTwist([1.0, 3.0])
    assert c.pos_cb((0.0, -0.99, 0.0)) == (0.0, 0.0) and c.done         # inside thresh: stop


def test_waypoints_from_path_and_planner_handoff():
    cells = [(0, 0), (1, 0), (2, 0), (3, 1), (4, 2), (5, 2)]
    w = waypoints_from_path(cells, min_spacing=0.25, origin=(-1.0, 2.0), resolution=0.1)
    assert w[0] == pytest.approx([-0.95, 2.05]) and w[-1] == pytest.approx([-0.45, 2.25])
    d = np.linalg.norm(np.diff(np.array(w), axis=0), axis=1)
    assert np.all(d >= 0.25 - 1e-12) and len(w) == 2
    assert waypoints_from_path([[0.0, 0.0], [0.01, 0.0], [1.0, 0.0]], min_spacing=0.1) == [[0.0, 0.0], [1.0, 0.0]]
    assert waypoints_from_path([]) == []

    class Stub(object):
        start = np.zeros(3)
        goal = np.zeros(3)
        thresh = 0.05
        uvec = np.array([[0.0, 0.0]])

        def initialize(self):
            pass

    c = Controller(mppi=Stub(), waypoints=None)
    c.pos_cb((0.0, 0.0, 0.0))
    c.set_waypoints(w)
    assert not c.parallel_park and c.init and c.idx == 0
    c.pos_cb((0.0, 0.0, 0.0))
    assert c.mppi.goal[:2] == pytest.approx(w[0]) and c.mppi.goal[2] == pytest.approx(np.arctan2(w[0][1], w[0][0]))
