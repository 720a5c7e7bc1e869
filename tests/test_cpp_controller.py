"""CPU: the C++ `mppi::Controller` (include/mppi.hpp) makes the same decisions as the Python mirror
`motion_planning_b200.Controller` -- which tests/test_controller_host.py pins to the unmodified reference node -- when
both drive the same deterministic stand-in engine over the same odometry."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from motion_planning_b200.controller import Controller
from oracle import ref_controller

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
LIBDIR = os.path.join(ROOT, "motion_planning_b200", "lib")
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


class StubMPPI(object):
    """Python twin of StubEngine in tests/cpp/controller_check.cpp, with the attribute surface Controller uses."""

    def __init__(self):
        self.start, self.goal, self.thresh = np.zeros(3), np.zeros(3), 0.05
        self.uvec = np.array([[0.0, 0.0]])
        self.steps = self.resets = 0

    def initialize(self):
        self.resets += 1
        self.uvec = np.array([[0.0, 0.0]])

    def get_path(self, x, g):
        self.steps += 1
        u = [0.5 + 0.25 * (g[0] - x[0]) - 0.125 * x[2], -0.75 + 0.5 * (g[1] - x[1]) + 0.0625 * g[2]]
        self.uvec = np.vstack([self.uvec, u])
        return x


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "controller_check")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "controller_check.cpp"),
           "-L" + LIBDIR, "-lmppi_b200", "-Wl,-rpath," + LIBDIR, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out


@pytest.mark.parametrize("waypoints", [[[0.3, 0.0], [0.3, 0.3], [0.0, 0.0]], []])
def test_cpp_controller_matches_python_mirror(exe, waypoints):
    rng = np.random.RandomState(1)
    stub = StubMPPI()
    node = Controller(mppi=stub, waypoints=waypoints)
    # a pose sequence that wanders through every branch: far from the goal, inside thresh of waypoints / the park goal
    targets = (waypoints or [[0.0, -1.0]]) * 3
    poses, quats = [], []
    p = np.array([0.1, -0.05, 0.3])
    for tgt in targets:
        for a in np.linspace(0.0, 1.0, 7):
            q = (1 - a) * p[:2] + a * np.array(tgt) + (0.0 if a == 1.0 else 1e-3 * rng.normal(size=2))
            poses.append([q[0], q[1], rng.uniform(-3.0, 3.0)])
        p = np.array([tgt[0], tgt[1], 0.0])
        poses.append([tgt[0] + 0.01, tgt[1] - 0.02, 0.5])       # a second sample inside thresh
    lines, expect = [], []
    for i, (x, y, th) in enumerate(poses):
        if i % 3 == 2:      # every third sample arrives as a quaternion, like nav_msgs/Odometry
            qx, qy, qz, qw = ref_controller.quaternion_from_yaw(th)
            lines.append("%.17g %.17g %.17g %.17g %.17g %.17g" % (x, y, qx, qy, qz, qw))
            vx, wz = node.pos_cb(ref_controller.make_odom(x, y, th))
        else:
            lines.append("%.17g %.17g %.17g" % (x, y, th))
            vx, wz = node.pos_cb((x, y, th))
        g = node.mppi.goal
        expect.append((node.idx, int(node.init), int(node.done), vx, wz, g[0], g[1], g[2], stub.steps, stub.resets))
    args = [str(v) for w in waypoints for v in w]
    r = subprocess.run([exe] + args, input="\n".join(lines) + "\n", capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = [l.split() for l in r.stdout.splitlines()]
    assert len(got) == len(expect)
    branches = set()
    for gline, e in zip(got, expect):
        assert (int(gline[0]), int(gline[1]), int(gline[2])) == e[:3]
        np.testing.assert_allclose([float(v) for v in gline[3:8]], e[3:8], rtol=0, atol=1e-12)
        assert (int(gline[8]), int(gline[9])) == e[8:]
        branches.add(e[:3][1:])
    assert stub.steps >= 3 and stub.resets >= (3 if waypoints else 1)
    if not waypoints:
        assert any(e[2] == 1 for e in expect)                   # parallel park reaches `done`


def test_cpp_path_tracking_matches_python_mirror(exe):
    """the look-ahead tracking variant (Controller.track_path / mppi::Controller::trackPath): same goals, same decisions."""
    path = [[0.0, 0.0], [0.5, 0.0], [0.5, 0.4], [0.9, 0.4], [0.9, 0.4], [1.0, 0.7]]       # with a repeated vertex
    L = 0.25
    rng = np.random.RandomState(3)
    stub = StubMPPI()
    node = Controller(mppi=stub)
    node.track_path(path, lookahead=L)
    # poses wandering along the polyline with some cross-track noise, then sitting at its end
    dense = []
    for a, b in zip(path[:-1], path[1:]):
        for u in np.linspace(0.0, 1.0, 6, endpoint=False):
            dense.append((1 - u) * np.array(a) + u * np.array(b))
    dense += [np.array(path[-1])] * 3
    lines, expect = [], []
    for q in dense:
        x, y, th = q[0] + 0.02 * rng.normal(), q[1] + 0.02 * rng.normal(), rng.uniform(-3, 3)
        lines.append("%.17g %.17g %.17g" % (x, y, th))
        vx, wz = node.pos_cb((x, y, th))
        g = node.mppi.goal
        expect.append((int(node.init), int(node.done), vx, wz, g[0], g[1], g[2], stub.steps, stub.resets))
    args = ["track", str(L)] + [str(v) for w in path for v in w]
    r = subprocess.run([exe] + args, input="\n".join(lines) + "\n", capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = [l.split() for l in r.stdout.splitlines()]
    assert len(got) == len(expect)
    for gline, e in zip(got, expect):
        assert (int(gline[1]), int(gline[2])) == e[:2]
        np.testing.assert_allclose([float(v) for v in gline[3:8]], e[2:7], rtol=0, atol=1e-12)
        assert (int(gline[8]), int(gline[9])) == e[7:]
    assert expect[-1][1] == 1 and stub.steps > 10 and stub.resets == 1      # reached the end, one initialise only


def test_cpp_host_rk4_matches_the_reference_integrator(tmp_path):
    """mppi::RK4 (the interface of control::RK4, control/include/control/rk4.hpp:19-62) over dd_dynamics against the oracle's
    rk4 WITHOUT the theta wrap (the C++ integrator has none, control/src/control/rk4.cpp:115-138): same stage order, so the
    trajectories agree to rounding; solve() returns floor(horizon / dt) states."""
    from oracle import mppi_oracle as orc
    out = str(tmp_path / "rk4_check")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "rk4_check.cpp"),
           "-L" + LIBDIR, "-lmppi_b200", "-Wl,-rpath," + LIBDIR, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    rng = np.random.RandomState(3)
    n, dt = 80, 1.0 / 64
    U = rng.uniform(-6.0, 6.0, size=(n, 2)) + np.array([-3.0, 3.0])   # turning left on average
    x0 = np.array([0.3, -0.2, 2.9])                       # close to pi: the trajectory crosses it, unwrapped
    text = "%r %r %d\n%r %r %r\n" % (dt, 1.0, n, *x0.tolist()) + "".join("%r %r\n" % (a, b) for a, b in U.tolist())
    r = subprocess.run([out], input=text, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    traj = np.array([[float(v) for v in line.split()] for line in r.stdout.splitlines()])
    assert traj.shape == (64, 3)                          # floor(1.0 / dt) steps although 80 controls were given
    x = x0.copy()
    step = orc.user_model_step(orc.dd_dynamics, "rk4", False)
    for i in range(64):
        x = step(x.reshape(3, 1), U[i].reshape(2, 1), dt)[:, 0]
        np.testing.assert_allclose(traj[i], x, rtol=1e-13, atol=1e-15)
    assert np.max(np.abs(traj[:, 2])) > np.pi             # no wrap in the C++ integrator
