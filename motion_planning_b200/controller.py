"""ROS-free mirror of the reference's `Controller` (control/src/mppi:296-389): the caller of the hot path.

The reference node subscribes to `odom`, runs one `MPPI.get_path` per message and publishes a `Twist`
on `cmd_vel`.  This class keeps that state machine line for line -- same attributes (`mppi`, `done`,
`parallel_park`, `waypoints`, `idx`, `init`, `state`), same `pos_cb` / `wheelsToTwist` -- with the two
ROS edges replaced by plain Python:

    rospy.Subscriber('odom', Odometry, pos_cb)      ->  call pos_cb(odom) yourself; `odom` is anything
                                                         shaped like nav_msgs/Odometry (pose.pose.position,
                                                         pose.pose.orientation) or a plain (x, y, theta)
    rospy.Publisher('cmd_vel', Twist).publish(tw)   ->  `publish(vx, wz)` callback; pos_cb also returns it
    rospy.get_param("waypoints")                    ->  constructor argument `waypoints` (falsy = parallel park)
    rospy.loginfo                                   ->  `log` callback (default: silent)

A rospy node keeps working unchanged: `Controller.pos_cb` accepts the Odometry message itself and the
node's `publish` callback fills the Twist.

Planner hand-off: the reference reads its waypoints from a yaml file
(control/config/waypoints.yaml:1) and never wires the planners in; `set_waypoints` /
`waypoints_from_path` accept what the planners produce -- the vertex list of `trace_path`
(global_planner/src/global_planner/heuristic.cpp:199-223) or the cell path of D* Lite
(incremental.cpp:291-336) -- and feed it to the same state machine.

Path tracking (the "tracking-cost variant"): with `track_path(path, lookahead)` the goal handed to
`MPPI.get_path` on every odometry sample is the LOOK-AHEAD point of the planner's polyline -- the point `lookahead` metres
of arc length beyond the robot's projection onto the path, heading along the path there -- so the quadratic goal cost of
control/src/mppi:165-171,180-184 becomes a path-tracking cost; the nominal control sequence is kept from step to step (the
goal moves continuously; no `initialize()` between samples as at a waypoint switch).  NEW: no reference counterpart.
"""
import math

import numpy as np

from .mppi import MPPI, WHEEL_BASE, WHEEL_RADIUS


def yaw_from_quaternion(x, y, z, w):
    """Yaw of tf.transformations.euler_from_quaternion([x, y, z, w])[2] (axes 'sxyz'; call site control/src/mppi:333-334).
    Same arithmetic as tf's route through quaternion_matrix, reduced to the two matrix entries the yaw needs: the
    quaternion is scaled by sqrt(2 / |q|^2) (so it need not be normalised), M10 = qx qy + qz qw, M00 = 1 - qy qy - qz qz."""
    nq = x * x + y * y + z * z + w * w
    if nq < 8.881784197001252e-16:              # tf: 4 * machine epsilon -> identity
        return 0.0
    s = math.sqrt(2.0 / nq)
    x, y, z, w = x * s, y * s, z * s, w * s
    return math.atan2(x * y + z * w, 1.0 - y * y - z * z)


def _pose_of(odom):
    """(x, y, theta) of an Odometry-shaped message (control/src/mppi:330-334) or of a plain triple."""
    pose = getattr(odom, "pose", None)
    if pose is None:
        x, y, theta = odom
        return float(x), float(y), float(theta)
    p, q = pose.pose.position, pose.pose.orientation
    return float(p.x), float(p.y), yaw_from_quaternion(q.x, q.y, q.z, q.w)


def waypoints_from_path(path, min_spacing=0.0, origin=(0.0, 0.0), resolution=None):
    """Waypoint list for `Controller` from a planner path.

    path        (n,2) vertices in metres (trace_path output) or, with `resolution`, integer grid cells
                (ix, iy) of a grid path, converted to cell centres with the map package's convention
                x = x_min + (ix + 0.5) res (map/src/map/grid.cpp:106-119).
    min_spacing drop vertices closer than this to the previously kept one (a cell-by-cell grid path would
                otherwise hand the controller goals already inside `thresh`); the last vertex is always kept.
    """
    p = np.asarray(path, dtype=np.float64).reshape(-1, 2)
    if resolution is not None:
        p = np.asarray(origin, dtype=np.float64)[None, :] + (p + 0.5) * float(resolution)
    if len(p) == 0:
        return []
    keep = [p[0]]
    for v in p[1:-1]:
        if np.linalg.norm(v - keep[-1]) >= min_spacing:
            keep.append(v)
    if len(p) > 1:
        if len(keep) > 1 and np.linalg.norm(p[-1] - keep[-1]) < min_spacing:
            keep[-1] = p[-1]
        else:
            keep.append(p[-1])
    return [[float(v[0]), float(v[1])] for v in keep]


def lookahead_goal(path, pos, lookahead, s_min=0.0):
    """Look-ahead point of a polyline.  path (n,2) vertices in driving order, pos (2,), lookahead in metres, s_min = arc
    length already covered (progress never moves backwards, so a path that crosses itself is followed in order).
    Returns (goal (3,) = (x, y, heading of the path at that point), s_proj = arc length of the robot's projection)."""
    p = np.asarray(path, dtype=np.float64).reshape(-1, 2)
    pos = np.asarray(pos, dtype=np.float64)[:2]
    if len(p) == 1:
        return np.array([p[0, 0], p[0, 1], 0.0]), 0.0
    seg = p[1:] - p[:-1]
    seglen = np.hypot(seg[:, 0], seg[:, 1])
    cum = np.concatenate([[0.0], np.cumsum(seglen)])
    # projection onto every segment, clamped to the segment and to the progress made so far
    best_d, s_proj = np.inf, s_min
    for i in range(len(seg)):
        if seglen[i] == 0.0 or cum[i + 1] < s_min:
            continue
        u = float(np.dot(pos - p[i], seg[i]) / (seglen[i] * seglen[i]))
        u = min(1.0, max(u, max(0.0, (s_min - cum[i]) / seglen[i])))
        q = p[i] + u * seg[i]
        d = float(np.hypot(pos[0] - q[0], pos[1] - q[1]))
        if d < best_d - 1e-12:
            best_d, s_proj = d, cum[i] + u * seglen[i]
    s_goal = min(s_proj + lookahead, cum[-1])
    i = int(np.searchsorted(cum, s_goal, side="right") - 1)
    i = min(max(i, 0), len(seg) - 1)
    while seglen[i] == 0.0 and i > 0:
        i -= 1
    u = (s_goal - cum[i]) / seglen[i] if seglen[i] > 0 else 0.0
    g = p[i] + u * seg[i]
    return np.array([g[0], g[1], math.atan2(seg[i, 1], seg[i, 0])]), s_proj


class Controller(object):
    def __init__(self, mppi=None, waypoints=None, publish=None, log=None):
        """control/src/mppi:297-319.  `mppi` defaults to `MPPI()` (K=10, T=100: what the node runs, :298)."""
        self.mppi = mppi if mppi is not None else MPPI()
        self._publish = publish
        self._log = log
        self.done = False
        if not waypoints:                                   # :305-309
            self.parallel_park = True
        else:
            self.parallel_park = False
            self.waypoints = [list(w) for w in waypoints]
        self.idx = 0
        self.init = True
        self._track = None
        self.state = self.mppi.start
        self.cmd = (0.0, 0.0)
        self._emit(0.0, 0.0)                                # :313-316: a zero Twist at start-up

    # ---- ROS edges ------------------------------------------------------------------------------
    def _emit(self, vx, wz):
        self.cmd = (vx, wz)
        if self._publish is not None:
            self._publish(vx, wz)

    def _info(self, msg):
        if self._log is not None:
            self._log(msg)

    def set_waypoints(self, waypoints):
        """Planner hand-off: replace the waypoint list; the next `pos_cb` re-initialises towards waypoints[0]."""
        if not waypoints:
            self.parallel_park = True
        else:
            self.parallel_park = False
            self.waypoints = [list(w) for w in waypoints]
        self.idx = 0
        self.init = True
        self.done = False
        self._track = None

    def track_path(self, path, lookahead=0.3):
        """Planner hand-off, tracking variant: follow the polyline `path` (trace_path vertices, metres) with a moving
        look-ahead goal instead of stopping at every vertex.  `pos_cb` then runs `_track_cb`."""
        self._track = np.asarray(path, dtype=np.float64).reshape(-1, 2).copy()
        self._lookahead = float(lookahead)
        self._progress = 0.0
        self.parallel_park = False
        self.waypoints = [list(self._track[-1])]
        self.idx = 0
        self.init = True
        self.done = False

    def _track_cb(self, x, y, theta):
        m = self.mppi
        m.start = np.array([x, y, theta])
        goal, self._progress = lookahead_goal(self._track, m.start, self._lookahead, self._progress)
        m.goal = goal
        end = self._track[-1]
        if self.init:                                       # first sample: zero nominal, like :344-354
            m.initialize()
            self.state = m.start
            self.init = False
            self._info("TRACKING {} vertices, look-ahead {} m".format(len(self._track), self._lookahead))
        elif np.linalg.norm(m.start[:2] - end) > m.thresh:
            self.state = m.get_path(m.start, m.goal)        # the hot path, goal = look-ahead point
            self.done = False
        else:
            self.done = True                                # end of the path reached: zero Twist, like :375
        u = m.uvec[-1, :] if not self.done else np.array([0.0, 0.0])
        vx, wz = self.wheelsToTwist(u)
        self._emit(vx, wz)
        return vx, wz

    # ---- reference methods ----------------------------------------------------------------------
    def wheelsToTwist(self, wheel_vels):
        """control/src/mppi:320-326."""
        ul = wheel_vels[0]
        ur = wheel_vels[1]
        vx = WHEEL_RADIUS * (ul + ur) / 2.0
        wz = WHEEL_RADIUS * (-ul + ur) / WHEEL_BASE
        return vx, wz

    def _goal_towards(self, wpt):
        """Goal = (wx, wy, heading from the current start to the waypoint), control/src/mppi:347-352."""
        theta = np.arctan2(wpt[1] - self.mppi.start[1], wpt[0] - self.mppi.start[0])
        return np.array([wpt[0], wpt[1], theta])

    def pos_cb(self, odom):
        """control/src/mppi:328-386: one odometry message -> one MPPI step -> one Twist (returned as (vx, wz))."""
        m = self.mppi
        x, y, theta = _pose_of(odom)
        if getattr(self, "_track", None) is not None:
            return self._track_cb(x, y, theta)
        m.start = np.array([x, y, theta])                   # :335
        if self.parallel_park:
            m.goal = np.array([0.0, -1.0, 0.0])             # :337
        if np.linalg.norm(m.start[:2] - m.goal[:2]) > m.thresh and not self.init:
            self.state = m.get_path(m.start, m.goal)        # :341 -- the hot path
            self.done = False
        elif self.init:                                     # :344-354
            m.initialize()
            if not self.parallel_park:
                m.goal = self._goal_towards(self.waypoints[self.idx])
                self.state = m.start
            self.init = False
            self._info("NEXT WAYPOINT: {}".format(m.goal[:2]))
        else:
            if not self.parallel_park:                      # :357-373: goal reached, next waypoint (cyclic)
                if self.idx + 1 >= len(self.waypoints):
                    self.idx = 0
                else:
                    self.idx += 1
                m.initialize()
                self._info("WAYPOINT REACHED: {} \n NEXT WAYPOINT: {}".format(m.goal[:2], self.waypoints[self.idx]))
                m.goal = self._goal_towards(self.waypoints[self.idx])
                self.state = m.start
            else:
                self.done = True                            # :375
        if not self.done:                                   # :378-381
            u = m.uvec[-1, :]
        else:
            u = np.array([0.0, 0.0])
        vx, wz = self.wheelsToTwist(u)
        self._emit(vx, wz)
        return vx, wz


class FakeDiffDrive(object):
    """Stand-in for the simulation half of control/launch/mppi_pentagon.launch:5-8,25-31 (rigid2d's fake
    encoders + odometer, which are not part of the reference repository): integrates the published Twist
    exactly over `dt` (constant-twist arc) and reports the pose as the next odometry."""

    def __init__(self, pose=(0.0, 0.0, 0.0), dt=0.01):
        self.pose = np.array(pose, dtype=np.float64)
        self.dt = float(dt)

    def step(self, vx, wz):
        x, y, th = self.pose
        a = wz * self.dt
        if abs(a) < 1e-12:
            x += vx * self.dt * math.cos(th)
            y += vx * self.dt * math.sin(th)
        else:
            r = vx / wz
            x += r * (math.sin(th + a) - math.sin(th))
            y -= r * (math.cos(th + a) - math.cos(th))
        th = math.atan2(math.sin(th + a), math.cos(th + a))
        self.pose = np.array([x, y, th])
        return self.pose


def run_closed_loop(controller, plant, n_callbacks):
    """Drive `controller` from `plant` for n odometry callbacks; returns the (n,3) poses fed in and the (n,2) twists out."""
    poses, twists = [], []
    for _ in range(int(n_callbacks)):
        pose = plant.pose.copy()
        vx, wz = controller.pos_cb(pose)
        plant.step(vx, wz)
        poses.append(pose)
        twists.append((vx, wz))
    return np.array(poses), np.array(twists)
