"""TEST INFRASTRUCTURE ONLY -- run the *unmodified* reference `Controller` (control/src/mppi:296-389) without ROS.

`ref_loader.load_reference()` executes the reference module with empty stand-ins for rospy / tf /
geometry_msgs / nav_msgs.  The MPPI class never touches them; `Controller` does, so this file fills
the stand-ins the loaded module holds (module globals `rospy`, `tf`, `Twist`) with the minimum a
node needs: a parameter server, a publisher that records, a no-op subscriber, and
`tf.transformations.euler_from_quaternion`.

tf is third-party and absent from this image (ROS 1 `tf` package, version unpinned by the reference:
control/package.xml lists it without a version).  `euler_from_quaternion` below restates its published
algorithm (transformations.py by C. Gohlke as shipped in ros/geometry: quaternion_matrix ->
euler_from_matrix, axes 'sxyz'), the general matrix route -- the product's closed-form yaw
(motion_planning_b200/controller.py: yaw_from_quaternion) is checked against it.

Only tests/ and tests/golden/make_golden.py import this file.
"""
import math
import types

import numpy as np

from . import ref_loader

_EPS = np.finfo(float).eps * 4.0


def quaternion_matrix(quaternion):
    """tf.transformations.quaternion_matrix for q = [x, y, z, w] (3x3 part)."""
    q = np.array(quaternion[:4], dtype=np.float64, copy=True)
    nq = np.dot(q, q)
    if nq < _EPS:
        return np.identity(3)
    q *= math.sqrt(2.0 / nq)
    q = np.outer(q, q)
    return np.array((
        (1.0 - q[1, 1] - q[2, 2], q[0, 1] - q[2, 3], q[0, 2] + q[1, 3]),
        (q[0, 1] + q[2, 3], 1.0 - q[0, 0] - q[2, 2], q[1, 2] - q[0, 3]),
        (q[0, 2] - q[1, 3], q[1, 2] + q[0, 3], 1.0 - q[0, 0] - q[1, 1])), dtype=np.float64)


def euler_from_quaternion(quaternion, axes="sxyz"):
    """tf.transformations.euler_from_quaternion, static xyz axes (the only form the reference uses, control/src/mppi:334)."""
    if axes != "sxyz":
        raise NotImplementedError(axes)
    M = quaternion_matrix(quaternion)
    cy = math.sqrt(M[0, 0] * M[0, 0] + M[1, 0] * M[1, 0])
    if cy > _EPS:
        ax = math.atan2(M[2, 1], M[2, 2])
        ay = math.atan2(-M[2, 0], cy)
        az = math.atan2(M[1, 0], M[0, 0])
    else:
        ax = math.atan2(-M[1, 2], M[1, 1])
        ay = math.atan2(-M[2, 0], cy)
        az = 0.0
    return ax, ay, az


def quaternion_from_yaw(theta):
    """[x, y, z, w] of a rotation about z (what a planar odometer publishes)."""
    return [0.0, 0.0, math.sin(0.5 * theta), math.cos(0.5 * theta)]


def make_odom(x, y, theta):
    """nav_msgs/Odometry-shaped object: pose.pose.position.{x,y,z}, pose.pose.orientation.{x,y,z,w}."""
    q = quaternion_from_yaw(theta)
    position = types.SimpleNamespace(x=float(x), y=float(y), z=0.0)
    orientation = types.SimpleNamespace(x=q[0], y=q[1], z=q[2], w=q[3])
    return types.SimpleNamespace(pose=types.SimpleNamespace(pose=types.SimpleNamespace(position=position, orientation=orientation)))


class _Twist(object):
    def __init__(self):
        self.linear = types.SimpleNamespace(x=0.0, y=0.0, z=0.0)
        self.angular = types.SimpleNamespace(x=0.0, y=0.0, z=0.0)


class _Publisher(object):
    def __init__(self, *args, **kwargs):
        self.sent = []

    def publish(self, tw):
        self.sent.append((float(tw.linear.x), float(tw.angular.z)))


def load_node(waypoints, mppi_kwargs=None):
    """Fresh reference module (=> np.random.seed(0), control/src/mppi:15) + a live `Controller`.

    waypoints   value of the ROS parameter "waypoints" (falsy = parallel park, control/src/mppi:305-309)
    mppi_kwargs if given, the node's `MPPI()` (control/src/mppi:298) is constructed with these keyword
                arguments instead of the defaults (K=10, T=100) -- the only deviation from the node as
                shipped, so that traces stay small.
    Returns (module, controller, log list)."""
    ref = ref_loader.load_reference()
    log = []
    ref.rospy.Subscriber = lambda *a, **k: None
    ref.rospy.Publisher = _Publisher
    ref.rospy.get_param = lambda name: {"waypoints": waypoints}[name]
    ref.rospy.loginfo = log.append
    ref.rospy.ROSInterruptException = type("ROSInterruptException", (Exception,), {})
    ref.tf.transformations = types.SimpleNamespace(euler_from_quaternion=euler_from_quaternion)
    ref.Twist = _Twist
    if mppi_kwargs:
        plain = ref.MPPI
        ref.MPPI = lambda: plain(**mppi_kwargs)
        try:
            node = ref.Controller()
        finally:
            ref.MPPI = plain
    else:
        node = ref.Controller()
    return ref, node, log


def run_node(node, plant_step, pose0, n_callbacks):
    """Drive the reference node in closed loop: pose -> pos_cb -> Twist -> plant_step(vx, wz) -> next pose.
    Returns poses (n,3), twists (n,2), the per-callback (idx, init, done) flags and the nominal sequence
    `latest_uvec` (n,2,T) the node's MPPI holds after each callback."""
    pose = np.array(pose0, dtype=np.float64)
    poses, twists, flags, nominal = [], [], [], []
    for _ in range(n_callbacks):
        poses.append(pose.copy())
        node.pos_cb(make_odom(*pose))
        vx, wz = node.tw_pub.sent[-1]
        twists.append((vx, wz))
        flags.append((node.idx, int(node.init), int(node.done)))
        nominal.append(np.array(node.mppi.latest_uvec, dtype=np.float64))
        pose = np.array(plant_step(vx, wz), dtype=np.float64)
    return np.array(poses), np.array(twists), np.array(flags), np.array(nominal)
