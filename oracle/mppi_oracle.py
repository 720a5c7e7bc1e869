"""TEST INFRASTRUCTURE ONLY -- float64 NumPy restatement of the reference MPPI step.

Restates ``/root/reference/control/src/mppi`` (Python 2/3 + NumPy, float64):
every function below cites the reference lines it follows.  It exists so that
(1) the CUDA path can be checked on the GPU box, where /root/reference is not
mounted, and (2) large-K cases finish in seconds (the reference loops over K in
Python, control/src/mppi:158-161).

PARITY PINNING: the reference ships no tests and no golden vectors
(SURVEY.md section 4).  This restatement is pinned against outputs of the
reference itself: tests/golden/*.npz are produced by tests/golden/make_golden.py
from the unmodified reference loaded by oracle/ref_loader.py, and
tests/test_oracle_golden.py checks this file against them (and, when the
reference is loadable, against a live run).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product never does.

Extensions that have NO reference counterpart (bicycle model, occupancy-grid
cost, total-cost weighting) are marked "NEW"; their parity is unpinned by the
reference and they are pinned only by this oracle.
"""
import numpy as np

# control/src/mppi:18-20
WHEEL_VEL_MAX = 6.35492
WHEEL_RADIUS = 0.033
WHEEL_BASE = 0.16

MODEL_DIFF_DRIVE = 0      # rk4 + dd_dynamics            control/src/mppi:23-30,39-54
MODEL_UNICYCLE_EULER = 1  # euler + unicycle_dynamics    control/src/mppi:33-36,57-58
MODEL_BICYCLE = 2         # NEW: kinematic bicycle, RK4 + wrap (no reference lines)
MODEL_USER = 3            # a caller's ODE through the reference's `model=` hook (control/src/mppi:62,66,154)

WEIGHT_COST_TO_GO = 0     # reference: per-t softmin on cost-to-go   control/src/mppi:175,187-196
WEIGHT_TOTAL_COST = 1     # NEW (north_star wording): one softmin on the total rollout cost


class Params(object):
    """All constants of one MPPI instance; defaults = the reference's hard-coded values."""

    def __init__(self, K=10, T=100, **kw):
        self.K = int(K)                                   # samples   control/src/mppi:62,64
        self.T = int(T)                                   # horizon   control/src/mppi:62,63
        self.dt = 1.0 / float(self.T)                     # control/src/mppi:67
        self.model = MODEL_DIFF_DRIVE
        self.weighting = WEIGHT_COST_TO_GO
        self.Q = np.array([1e3, 1e3, 0.0])                # diag, control/src/mppi:69
        self.R = np.array([[1.0, 0.0], [0.0, 1.0]])       # control/src/mppi:71
        self.P1 = np.array([1e3, 1e3, 1e3])               # diag, control/src/mppi:73
        self.sig = np.array([[0.9, 0.0], [0.0, 0.9]])     # control/src/mppi:88
        self.noise_std = np.array([0.9, 0.9])             # = sig[0,0] for both rows, control/src/mppi:144-146
        self.lam = 1e-3                                   # control/src/mppi:89
        self.u_max = np.array([WHEEL_VEL_MAX, WHEEL_VEL_MAX])   # control/src/mppi:151-152
        self.wheel_radius = WHEEL_RADIUS
        self.wheel_base = WHEEL_BASE
        self.eps_floor = 1e-8                             # control/src/mppi:193
        # NEW: occupancy grid term
        self.grid = None          # int8 (H, W) row-major, idx = ix + iy*W   map/src/map/grid.cpp:251-266
        self.grid_res = 1.0
        self.grid_origin = np.array([0.0, 0.0])           # map_min           map/src/viz_grid.cpp:112-129
        self.w_obs = 0.0
        # MODEL_USER: the integrator-step functor the reference takes as MPPI(model=...) (control/src/mppi:62,66), built from
        # an ODE right-hand side f(x (3,N), u (2,N)) -> (3,N) like the reference's own rk4 / euler are (:39-58)
        self.user_ode = None
        self.user_integrator = "rk4"      # "rk4" (:39-50) | "euler" (:57-58)
        self.user_wrap = True             # theta wrap of :52-53
        # optional cost functors replacing get_cost (:180-184) and the terminal cost (:165-171):
        #   user_running_cost(st (3,N), goal (3,), u_nom (2,), eps_t (2,N), t) -> (N,);  user_terminal_cost(st, goal) -> (N,)
        self.user_running_cost = None
        self.user_terminal_cost = None
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


# --------------------------------------------------------------------------- dynamics

def dd_dynamics(x, u, r=WHEEL_RADIUS, L=WHEEL_BASE):
    """control/src/mppi:23-30."""
    return np.array([
        (r / 2.0) * np.cos(x[2, :]) * (u[0, :] + u[1, :]),
        (r / 2.0) * np.sin(x[2, :]) * (u[0, :] + u[1, :]),
        (r / L) * (u[1, :] - u[0, :]),
    ])


def unicycle_dynamics(x, u):
    """control/src/mppi:33-36."""
    return np.array([np.cos(x[2, :]) * u[0, :], np.sin(x[2, :]) * u[0, :], u[1, :]])


def bicycle_dynamics(x, u, L=WHEEL_BASE):
    """NEW. Kinematic bicycle: u = (v, delta); xdot = v cos th, ydot = v sin th, thdot = v tan(delta)/L."""
    return np.array([np.cos(x[2, :]) * u[0, :], np.sin(x[2, :]) * u[0, :], u[0, :] * np.tan(u[1, :]) / L])


def _wrap(th):
    """control/src/mppi:52-53: theta - (ceil((theta+pi)/(2pi)) - 1) 2pi  -> (-pi, pi]."""
    return th - (np.ceil((th + np.pi) / (2.0 * np.pi)) - 1.0) * 2.0 * np.pi


def rk4(x0, u, dt, f=dd_dynamics):
    """control/src/mppi:39-54 (u held constant over the step; theta wrapped afterwards)."""
    k1 = dt * f(x0, u)
    k2 = dt * f(x0 + k1 / 2, u)
    k3 = dt * f(x0 + k2 / 2, u)
    k4 = dt * f(x0 + k3, u)
    xnew = x0 + (1.0 / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
    xnew[2, :] = _wrap(xnew[2, :])
    return xnew


def euler(x0, u, dt):
    """control/src/mppi:57-58 (no wrap)."""
    return x0 + dt * unicycle_dynamics(x0, u)


def model_step(p, x, u):
    if p.model == MODEL_DIFF_DRIVE:
        return rk4(x, u, p.dt, lambda a, b: dd_dynamics(a, b, p.wheel_radius, p.wheel_base))
    if p.model == MODEL_UNICYCLE_EULER:
        return euler(x, u, p.dt)
    if p.model == MODEL_BICYCLE:
        return rk4(x, u, p.dt, lambda a, b: bicycle_dynamics(a, b, p.wheel_base))
    if p.model == MODEL_USER:
        return user_model_step(p.user_ode, p.user_integrator, p.user_wrap)(x, u, p.dt)
    raise ValueError("model")


def user_model_step(f, integrator="rk4", wrap=True):
    """An integrator-step functor `model(states, u, dt)` for the reference's `MPPI(model=...)` hook (control/src/mppi:62,66,
    154), built from an ODE right-hand side exactly as the reference builds its own: rk4 (:39-50) or euler (:57-58), with or
    without the theta wrap (:52-53)."""
    def model(x0, u, dt):
        if integrator == "euler":
            xnew = x0 + dt * f(x0, u)
        else:
            k1 = dt * f(x0, u)
            k2 = dt * f(x0 + k1 / 2, u)
            k3 = dt * f(x0 + k2 / 2, u)
            k4 = dt * f(x0 + k3, u)
            xnew = x0 + (1.0 / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
        if wrap:
            xnew[2, :] = _wrap(xnew[2, :])
        return xnew
    return model


# --------------------------------------------------------------------------- grid (NEW)

def grid_cost(p, st):
    """NEW. value/100 * w_obs; outside the map counts as occupied (100).

    cell = (floor((x-x_min)/res), floor((y-y_min)/res)), idx = ix + iy*W
    (map/src/map/grid.cpp:71-104 world2grid, :251-266 grid2rowmajor); values 0/50/100
    (map/src/map/grid.cpp:126-144); out of bounds: the reference throws (grid.cpp:93-96).
    """
    if p.grid is None or p.w_obs == 0.0:
        return 0.0
    H, W = p.grid.shape
    ix = np.floor((st[0, :] - p.grid_origin[0]) / p.grid_res)
    iy = np.floor((st[1, :] - p.grid_origin[1]) / p.grid_res)
    inside = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
    ixc = np.clip(ix, 0, W - 1).astype(np.int64)
    iyc = np.clip(iy, 0, H - 1).astype(np.int64)
    val = np.where(inside, p.grid[iyc, ixc].astype(np.float64), 100.0)
    return p.w_obs * val / 100.0


# --------------------------------------------------------------------------- hot loop 1

def get_cost2go(p, state, uvec, goal, eps):
    """control/src/mppi:127-178 with the noise passed in.

    eps: (T, 2, K) float64 -- eps[t] is what ``np.random.normal(0, sig[0,0], size=(2,K))``
    returned at step t (control/src/mppi:143-146).  Returns V (T, K).
    """
    K, T = p.K, p.T
    states = np.tile(np.asarray(state, dtype=np.float64), (K, 1)).T            # :131
    goal = np.asarray(goal, dtype=np.float64)
    cost2go = np.empty((T, K))
    for t in range(T):
        u_samp = np.tile(uvec[:, t], (K, 1)).T + eps[t]                         # :147-149
        u_samp[0, :] = np.clip(u_samp[0, :], -p.u_max[0], p.u_max[0])           # :151
        u_samp[1, :] = np.clip(u_samp[1, :], -p.u_max[1], p.u_max[1])           # :152
        st = model_step(p, states, u_samp)                                      # :154
        states = st                                                             # :159
        d = st - goal[:, None]
        u = uvec[:, t]
        # get_cost, control/src/mppi:180-184 -- x is the post-step state, u the NOMINAL control
        if p.user_running_cost is not None:
            cost2go[t] = p.user_running_cost(st, goal, u, eps[t], t)
        else:
            q = p.Q[0] * d[0] * d[0] + p.Q[1] * d[1] * d[1] + p.Q[2] * d[2] * d[2]
            cost2go[t] = 0.5 * (q + u.dot(p.R).dot(u)) + p.lam * (u.dot(p.sig)).dot(eps[t])
        cost2go[t] += grid_cost(p, st)
    d = st - goal[:, None]                                                      # :165-171 (theta NOT wrapped)
    if p.user_terminal_cost is not None:
        cost2go[-1] += p.user_terminal_cost(st, goal)
    else:
        cost2go[-1] += p.P1[0] * d[0] * d[0] + p.P1[1] * d[1] * d[1] + p.P1[2] * d[2] * d[2]
    V = np.flip(np.cumsum(np.flip(cost2go, 0), axis=0), 0)                      # :175
    if p.weighting == WEIGHT_TOTAL_COST:
        V = np.tile(V[0], (T, 1))
    return V


# --------------------------------------------------------------------------- SavGol (third party)

def savgol_matrix(T):
    """The fixed linear map applied by ``scipy.signal.savgol_filter(U, T-1, 3, axis=1)``
    (call site control/src/mppi:202; SciPy is an un-pinned third-party dependency).

    Published algorithm (scipy/signal/_savitzky_golay.py: savgol_filter, mode='interp'):
    window w = T-1, half h = w//2.  Interior outputs (h .. T-h-1) are the centred
    least-squares cubic evaluated at the window centre; the first h outputs evaluate the
    cubic fitted to the FIRST w samples, the last h outputs the cubic fitted to the LAST w
    samples (_fit_edges_polyfit).  Returns S (T, T) with  filtered = U @ S.T  row-wise.
    """
    w = T - 1
    if w % 2 != 1 or w < 5:
        raise ValueError("savgol window T-1 must be odd and > polyorder+1 (T even, T >= 6)")
    h = w // 2
    z = np.arange(w, dtype=np.float64) - h                 # centred abscissae of a window
    A = np.vander(z, 4, increasing=True)                   # (w, 4)
    P = np.linalg.pinv(A)                                  # (4, w): coefficients = P @ samples

    def rows(at):                                          # hat-matrix rows evaluating the fit at z=at
        return np.vander(np.asarray(at, dtype=np.float64), 4, increasing=True) @ P

    S = np.zeros((T, T))
    S[0:h + 1, 0:w] = rows(np.arange(0, h + 1) - h)        # outputs 0..h   from window [0, w)
    S[T - h - 1:T, 1:T] = rows(np.arange(T - h - 1, T) - 1 - h)   # outputs T-h-1..T-1 from window [1, T)
    return S


def update_action(p, uvec, eps, V):
    """control/src/mppi:186-208.  uvec (2,T), eps (T,2,K), V (T,K); returns the new (2,T)."""
    uvec = np.array(uvec, dtype=np.float64)
    V = np.array(V, dtype=np.float64)
    for t in range(p.T):
        V[t] -= np.amin(V[t])                                                   # :189
        omg = np.exp(-V[t] / p.lam) + p.eps_floor                               # :193
        omg /= np.sum(omg)                                                      # :195
        uvec[:, t] += np.dot(eps[t], omg)                                       # :196
    pre_filter = uvec.copy()
    uvec[0, :] = np.clip(uvec[0, :], -p.u_max[0], p.u_max[0])                   # :198
    uvec[1, :] = np.clip(uvec[1, :], -p.u_max[1], p.u_max[1])                   # :199
    uvec = uvec @ savgol_matrix(p.T).T                                          # :202
    uvec[0, :] = np.clip(uvec[0, :], -p.u_max[0], p.u_max[0])                   # :205
    uvec[1, :] = np.clip(uvec[1, :], -p.u_max[1], p.u_max[1])                   # :206
    return uvec, pre_filter


def perform_action(p, state, uvec):
    """control/src/mppi:210-213."""
    st = np.tile(np.asarray(state, dtype=np.float64), (2, 1)).T
    u = np.tile(uvec[:, 0], (2, 1)).T
    return model_step(p, st, u)[:, 0]


def step(p, state, goal, U, eps):
    """One ``MPPI.get_path`` (control/src/mppi:85-102) on explicit noise.

    Returns dict: V (T,K), U_pre_filter, U_new (before the shift), u0 (= uvec[-1]),
    x_next (returned state), U_shift (latest_uvec after the receding-horizon shift).
    """
    V = get_cost2go(p, state, U, goal, eps)                                     # :90
    U_new, pre = update_action(p, U, eps, V)                                    # :92
    x_next = perform_action(p, state, U_new)                                    # :94
    U_shift = U_new.copy()
    U_shift[:, :-1] = U_new[:, 1:]                                              # :100
    U_shift[:, -1] = 0.0                                                        # :101 (uvec_init is zeros)
    return dict(V=V, U_pre_filter=pre, U_new=U_new, u0=U_new[:, 0].copy(), x_next=x_next, U_shift=U_shift)


def draw_reference_noise(p, rng=np.random):
    """The noise exactly as the reference draws it: T calls of normal(0, sig[0,0], (2,K))
    (control/src/mppi:143-146) on the legacy global MT19937 stream."""
    return np.stack([rng.normal(0, p.sig[0, 0], size=(2, p.K)) for _ in range(p.T)])


# --------------------------------------------------------------------------- diagnostics

def softmin_gaps(V):
    """gap(best, 2nd best) per t -- the conditioning indicator of SURVEY appendix C."""
    s = np.sort(V, axis=1)
    return s[:, 1] - s[:, 0]


# The occupancy-grid INPUT FORMAT (Grid::build_map / occupancy_grid of the map package) is restated in oracle/map_grid.py,
# pinned cell by cell against the reference's own map sources compiled by oracle/build_map_ref.py.
