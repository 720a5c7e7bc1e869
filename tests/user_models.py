"""Test fixtures: caller-supplied models as (CUDA text for the engine, NumPy twin for the oracle / the reference's `model=` hook).
The NumPy functions have the signature of the reference's own dynamics functions (control/src/mppi:23-36): f(x (3,N), u (2,N))."""
import numpy as np

# (1) the reference's diff-drive ODE itself (dd_dynamics, control/src/mppi:23-30), written as a user functor
DD_CUDA = """
template <typename R> __device__ void mppi_user_ode(const R x[3], const R u[2], R xdot[3]) {
  const R r = R(0.033), L = R(0.16);
  xdot[0] = (r / R(2.0)) * cos(x[2]) * (u[0] + u[1]);
  xdot[1] = (r / R(2.0)) * sin(x[2]) * (u[0] + u[1]);
  xdot[2] = (r / L) * (u[1] - u[0]);
}
"""


def dd_numpy(x, u):
    r, L = 0.033, 0.16
    return np.array([(r / 2.0) * np.cos(x[2, :]) * (u[0, :] + u[1, :]), (r / 2.0) * np.sin(x[2, :]) * (u[0, :] + u[1, :]),
                     (r / L) * (u[1, :] - u[0, :])])


# (2) a model NONE of the built-in kernels can express: speed and yaw rate depend on the state (a diff-drive on a surface whose
#     traction varies with x and that pulls the heading with y)
SKID_CUDA = """
template <typename R> __device__ void mppi_user_ode(const R x[3], const R u[2], R xdot[3]) {
  const R r = R(0.033), L = R(0.16);
  const R grip = R(1.0) - R(0.2) * tanh(x[0]);
  xdot[0] = (r / R(2.0)) * cos(x[2]) * (u[0] + u[1]) * grip;
  xdot[1] = (r / R(2.0)) * sin(x[2]) * (u[0] + u[1]) * grip;
  xdot[2] = (r / L) * (u[1] - u[0]) + R(0.4) * sin(R(2.0) * x[1]);
}
"""


def skid_numpy(x, u):
    r, L = 0.033, 0.16
    grip = 1.0 - 0.2 * np.tanh(x[0, :])
    return np.array([(r / 2.0) * np.cos(x[2, :]) * (u[0, :] + u[1, :]) * grip, (r / 2.0) * np.sin(x[2, :]) * (u[0, :] + u[1, :]) * grip,
                     (r / L) * (u[1, :] - u[0, :]) + 0.4 * np.sin(2.0 * x[1, :])])


# (3) a cost functor: the reference's running cost (control/src/mppi:180-184, lam = 1e-3, sig = 0.9 I, Q = diag(1e3, 1e3, 0), R = I)
#     plus a repulsive potential around (0.3, -0.2); the terminal cost stays the reference's (P1 = 1e3 I, :165-171)
COST_CUDA = """
template <typename R> __device__ R mppi_user_running_cost(const R x[3], const R g[3], const R u[2], const R eps[2], int t) {
  const R dx = x[0] - g[0], dy = x[1] - g[1];
  const R ox = x[0] - R(0.3), oy = x[1] + R(0.2);
  return R(0.5) * (R(1000.0) * dx * dx + R(1000.0) * dy * dy + u[0] * u[0] + u[1] * u[1])
       + R(0.001) * (R(0.9) * u[0] * eps[0] + R(0.9) * u[1] * eps[1]) + R(2.0) / (R(0.05) + ox * ox + oy * oy);
}
template <typename R> __device__ R mppi_user_terminal_cost(const R x[3], const R g[3]) {
  const R dx = x[0] - g[0], dy = x[1] - g[1], dt = x[2] - g[2];
  return R(1000.0) * (dx * dx + dy * dy + dt * dt);
}
"""


def running_cost_numpy(st, goal, u, eps_t, t):
    dx, dy = st[0] - goal[0], st[1] - goal[1]
    ox, oy = st[0] - 0.3, st[1] + 0.2
    return 0.5 * (1000.0 * dx * dx + 1000.0 * dy * dy + u[0] * u[0] + u[1] * u[1]) \
        + 0.001 * (0.9 * u[0] * eps_t[0] + 0.9 * u[1] * eps_t[1]) + 2.0 / (0.05 + ox * ox + oy * oy)


def terminal_cost_numpy(st, goal):
    d = st - np.asarray(goal)[:, None]
    return 1000.0 * (d[0] * d[0] + d[1] * d[1] + d[2] * d[2])


# (4) KINEMATIC functors (mppi_user_model.kind 1): speed and yaw rate from the controls.  The reference's diff-drive (:23-30) ...
DD_KIN_CUDA = """
template <typename R> __device__ void mppi_user_speed_yaw(const R u[2], R* speed, R* yaw_rate) {
  *speed = R(0.5 * 0.033) * (u[0] + u[1]);
  *yaw_rate = R(0.033 / 0.16) * (u[1] - u[0]);
}
"""
DD_KIN_BOUNDS = dict(speed_max=0.033 * 6.35492, yaw_rate_max=0.033 / 0.16 * 2 * 6.35492)

# ... and a vehicle none of the built-in models is: a tracked base whose tracks slip more the harder they are driven apart
SLIP_KIN_CUDA = """
template <typename R> __device__ void mppi_user_speed_yaw(const R u[2], R* speed, R* yaw_rate) {
  const R r = R(0.033), L = R(0.16);
  const R d = u[1] - u[0];
  const R slip = R(1.0) / (R(1.0) + R(0.02) * d * d);
  *speed = R(0.5) * r * (u[0] + u[1]) * (R(0.6) + R(0.4) * slip);
  *yaw_rate = r / L * d * slip;
}
"""
# |speed| <= r * u_max; |yaw| = (r / L) |d| / (1 + 0.02 d^2) is largest at |d| = 1 / sqrt(0.02)
SLIP_KIN_BOUNDS = dict(speed_max=0.033 * 6.35492, yaw_rate_max=0.033 / 0.16 * 0.5 / np.sqrt(0.02))


def slip_numpy(x, u):
    r, L = 0.033, 0.16
    d = u[1, :] - u[0, :]
    slip = 1.0 / (1.0 + 0.02 * d * d)
    s = 0.5 * r * (u[0, :] + u[1, :]) * (0.6 + 0.4 * slip)
    return np.array([s * np.cos(x[2, :]), s * np.sin(x[2, :]), r / L * d * slip])
