"""Host-side mirror of the reference's `MPPI` class (control/src/mppi:61-213) over the C ABI.

Same constructor, attributes, method names, argument meaning and return types as the reference, so
the ROS `Controller` (control/src/mppi:296-389) can do `self.mppi = MPPI()` unchanged:

    reference (NumPy, CPU)                         this class (libmppi_b200.so, B200)
    -------------------------------------------    -------------------------------------------
    MPPI(model=rk4, horizon=100, samples=10,       same + keyword-only engine options
         thresh=0.05)                 :62-77
    initialize()                      :79-83       mppi_reset
    get_path(state, goal, sig, lam)   :85-102      mppi_step  (2 kernel launches, no copy: engine.cu)
    solve_path(start, goal, sig, lam) :104-125     loop over get_path
    get_cost2go(state,uvec,goal,lam,sig) :127-178  mppi_cost_to_go (Philox noise drawn on device)
    update_action(uvec,eps,V,sig,lam) :186-208     mppi_update_action
    perform_action(state, uvec)       :210-213     mppi_perform_action
    attrs horizon samples dt Q R P1 thresh start goal uvec_init latest_uvec uvec path fin_time

All arrays are float64 ndarrays owned by the caller (fresh copies), like the reference.  Errors: the
reference raises nothing on the hot path (bad input propagates NaN); here a non-finite update raises
MppiError(MPPI_ERR_NONFINITE) instead of silently publishing NaN wheel speeds.
"""
import copy
import ctypes as C
import time

import numpy as np

from . import _capi

# robot constants, control/src/mppi:18-20
WHEEL_VEL_MAX = 6.35492
WHEEL_RADIUS = 0.033
WHEEL_BASE = 0.16


class _Model(object):
    """Integrator-step functor tag (the reference passes the function object `rk4`, :62).
    Calling it runs that model on the device: model(states (3,N), u (2,N), dt) -> (3,N)."""

    def __init__(self, name, model_id):
        self.__name__ = name
        self.model_id = model_id

    def _create_engine(self, lib, p, h):
        _capi.check(lib.mppi_create(C.byref(p), C.byref(h)), "mppi_create")

    def __call__(self, x0, u, dt):
        x0 = _capi.f64(x0)
        u = _capi.f64(u)
        n = x0.shape[1]
        p = _capi.MppiParams()
        lib = _capi.load()
        _capi.check(lib.mppi_default_params(C.byref(p)), "mppi_default_params")
        p.K, p.T, p.model, p.dt = 1, 6, self.model_id, float(dt)
        p.precision = _capi.PRECISION_F64
        h = C.c_void_p()
        self._create_engine(lib, p, h)
        try:
            out = np.empty((3, n))
            _capi.check(lib.mppi_model_step(h, _capi.dptr(x0), _capi.dptr(u), n, _capi.dptr(out)), "mppi_model_step")
        finally:
            lib.mppi_destroy(h)
        return out

    def __repr__(self):
        return "<mppi model %s>" % self.__name__


class UserModel(_Model):
    """A caller-supplied model: the integrator-step functor the reference takes as `MPPI(model=...)` (control/src/mppi:62,66),
    given as CUDA text and compiled for sm_100a at run time (mppi_create_user, include/mppi_b200.h).

        ode_source    template <typename R> __device__ void mppi_user_ode(const R x[3], const R u[2], R xdot[3]) {...}
                      -- the signature of the reference C++ library's ODE functor, control/include/control/rk4.hpp:32,58
        integrator    "rk4" (control held over the step, control/src/mppi:39-50) or "euler" (:57-58)
        wrap_theta    wrap theta into (-pi, pi] after every step (:52-53)
        cost_source   optional: mppi_user_running_cost<R>(x, goal, u_nom, eps, t) and mppi_user_terminal_cost<R>(x, goal)
                      replacing get_cost (:180-184) and the terminal cost (:165-171)

    Engines with a user model run with precision 'f64' (default) or 'f32'.  See KinematicModel for functors that also run
    the headline precision 'mixed'."""

    kind = 0
    speed_max = 0.0
    yaw_rate_max = 0.0

    def __init__(self, ode_source, name="user_model", integrator="rk4", wrap_theta=True, cost_source=None):
        _Model.__init__(self, name, _capi.MODEL_USER)
        self.source = ode_source + ("\n" + cost_source if cost_source else "")
        self.integrator = {"rk4": 0, "euler": 1}[integrator]
        self.wrap_theta = bool(wrap_theta)
        self.has_cost = cost_source is not None

    def _spec(self):
        self._src_bytes = self.source.encode("utf-8")       # kept alive for the duration of the call
        return _capi.MppiUserModel(self._src_bytes, self.integrator, int(self.wrap_theta), int(self.has_cost), self.kind,
                                   float(self.speed_max), float(self.yaw_rate_max))

    def _screenable(self):
        """can the engine run precision 'mixed' for this model?"""
        return self.kind == 1 and not self.has_cost and self.speed_max > 0

    def check(self):
        """compile only (no GPU needed); raises MppiError with the compiler's log"""
        um = self._spec()
        _capi.check(_capi.load().mppi_check_user_model(C.byref(um)), "mppi_check_user_model")

    def _create_engine(self, lib, p, h):
        um = self._spec()
        _capi.check(lib.mppi_create_user(C.byref(p), C.byref(um), C.byref(h)), "mppi_create_user")


class KinematicModel(UserModel):
    """A caller-supplied KINEMATIC model (mppi_user_model.kind 1): forward speed and yaw rate as functions of the controls,

        source        template <typename R> __device__ void mppi_user_speed_yaw(const R u[2], R* speed, R* yaw_rate) {...}

    for xdot = speed cos(theta), ydot = speed sin(theta), thetadot = yaw_rate -- the family dd_dynamics and unicycle_dynamics
    (control/src/mppi:23-36) belong to.  The functor is dropped into the built-in kernels, so every precision works, 'mixed'
    (the default, as for the built-in models) included; 'mixed' needs speed_max / yaw_rate_max, bounds of |speed| and
    |yaw_rate| over the clipped controls (they size the fp32 screening window; a bound that is too small costs fp64 re-runs,
    not correctness).  integrator "rk4" wraps theta after the step (:39-54), "euler" does not (:57-58)."""

    kind = 1

    def __init__(self, source, name="kinematic_model", integrator="rk4", speed_max=0.0, yaw_rate_max=0.0, cost_source=None):
        UserModel.__init__(self, source, name=name, integrator=integrator, wrap_theta=(integrator == "rk4"), cost_source=cost_source)
        self.speed_max = speed_max
        self.yaw_rate_max = yaw_rate_max


rk4 = _Model("rk4", _capi.MODEL_DIFF_DRIVE)                 # control/src/mppi:39-54 (+ dd_dynamics :23-30)
euler = _Model("euler", _capi.MODEL_UNICYCLE_EULER)         # control/src/mppi:57-58 (+ unicycle_dynamics :33-36)
bicycle_rk4 = _Model("bicycle_rk4", _capi.MODEL_BICYCLE)    # NEW (BASELINE.json config 3)

_PRECISIONS = {"f32": _capi.PRECISION_F32, "f64": _capi.PRECISION_F64, "mixed": _capi.PRECISION_MIXED}
_WEIGHTINGS = {"cost_to_go": _capi.WEIGHT_COST_TO_GO, "total_cost": _capi.WEIGHT_TOTAL_COST}


def _diag3(M, name):
    M = np.asarray(M, dtype=np.float64)
    if M.shape == (3,):
        return M
    if M.shape != (3, 3) or np.any(M - np.diag(np.diag(M)) != 0):
        raise NotImplementedError("%s must be diagonal (the reference's is, control/src/mppi:69-73)" % name)
    return np.diag(M).copy()


class _Log(object):
    """Append-only (n, d) float64 log with amortised O(1) append; `.array` is the (n, d) view.
    (The reference re-concatenates `path` / `uvec` every step, control/src/mppi:95-97: O(n) per step.)"""

    def __init__(self, first_row):
        first_row = np.asarray(first_row, dtype=np.float64).reshape(1, -1)
        self._buf = np.empty((64, first_row.shape[1]))
        self._buf[0] = first_row[0]
        self._n = 1

    @classmethod
    def from_array(cls, a):
        a = np.asarray(a, dtype=np.float64)
        if a.ndim != 2 or a.shape[0] < 1:
            raise ValueError("expected an (n, d) array with n >= 1")
        log = cls(a[0])
        for row in a[1:]:
            log.append(row)
        return log

    def append(self, row):
        if self._n == self._buf.shape[0]:
            grown = np.empty((2 * self._n, self._buf.shape[1]))
            grown[:self._n] = self._buf
            self._buf = grown
        self._buf[self._n] = row
        self._n += 1

    @property
    def array(self):
        return self._buf[:self._n]


class MPPI(object):
    def __init__(self, model=rk4, horizon=100, samples=10, thresh=0.05, **engine):
        """Reference signature (control/src/mppi:62) + keyword-only engine options:
        precision='mixed'|'f32'|'f64', weighting='cost_to_go'|'total_cost', seed, device,
        u_max, noise_std, wheel_radius, wheel_base, k_offset, k_total, world_size, rank, stream,
        refine_margin."""
        if not isinstance(model, _Model):
            name = getattr(model, "__name__", "")
            model = {"rk4": rk4, "euler": euler}.get(name)
            if model is None:
                raise TypeError("model must be one of rk4 / euler / bicycle_rk4 or a UserModel (CUDA text compiled at run time); "
                                "arbitrary Python callables cannot run inside the rollout kernel")
        self.horizon = int(horizon)                                             # :63
        self.samples = int(samples)                                             # :64
        self.uvec_init = np.zeros((2, self.horizon))                            # :65
        self.model = model                                                      # :66
        self.dt = 1.0 / float(horizon)                                          # :67
        self.Q = np.array([[1e3, 0.0, 0.0], [0.0, 1e3, 0.0], [0.0, 0.0, 0.0]])  # :69
        self.R = np.array([[1.0, 0.0], [0.0, 1.0]])                             # :71
        self.P1 = np.array([[1e3, 0.0, 0.0], [0.0, 1e3, 0.0], [0.0, 0.0, 1e3]])  # :73
        self.thresh = thresh                                                    # :74
        self.start = np.array([0.0, 0.0, 0.0])                                  # :75
        self.goal = np.array([0.0, 0.0, 0.0])                                   # :76
        self._engine_opts = dict(engine)
        self._lib = _capi.load()
        self._h = None
        self._sig = np.array([[.9, 0.0], [0.0, .9]])
        self._lam = .001
        self._cost_key = None
        self._create()
        self.initialize()

    # ---- engine plumbing -------------------------------------------------------------------
    def _create(self):
        o = dict(self._engine_opts)
        p = _capi.MppiParams()
        _capi.check(self._lib.mppi_default_params(C.byref(p)), "mppi_default_params")
        p.K, p.T = self.samples, self.horizon
        p.model = self.model.model_id
        p.dt = self.dt
        p.precision = _PRECISIONS[o.pop("precision", "f64" if (isinstance(self.model, UserModel) and not self.model._screenable()) else "mixed")]
        p.weighting = _WEIGHTINGS[o.pop("weighting", "cost_to_go")]
        p.seed = int(o.pop("seed", 0))
        p.device = int(o.pop("device", 0))
        q, p1 = _diag3(self.Q, "Q"), _diag3(self.P1, "P1")
        for i in range(3):
            p.q[i], p.p1[i] = q[i], p1[i]
        for i, v in enumerate(np.asarray(self.R, dtype=np.float64).reshape(4)):
            p.r[i] = v
        um = np.broadcast_to(np.asarray(o.pop("u_max", WHEEL_VEL_MAX), dtype=np.float64), (2,))
        p.u_max[0], p.u_max[1] = um[0], um[1]
        ns = o.pop("noise_std", None)
        self._noise_std_override = None if ns is None else np.broadcast_to(np.asarray(ns, dtype=np.float64), (2,)).copy()
        p.wheel_radius = float(o.pop("wheel_radius", WHEEL_RADIUS))
        p.wheel_base = float(o.pop("wheel_base", WHEEL_BASE))
        p.k_offset = int(o.pop("k_offset", 0))
        p.k_total = int(o.pop("k_total", 0))
        p.world_size = int(o.pop("world_size", 1))
        p.rank = int(o.pop("rank", 0))
        p.stream = o.pop("stream", None)
        p.refine_margin = float(o.pop("refine_margin", 0.0))
        if o:
            raise TypeError("unknown engine options: %s" % sorted(o))
        h = C.c_void_p()
        self.model._create_engine(self._lib, p, h)
        self._h = h
        self._cost_key = (tuple(q), tuple(p1), tuple(np.asarray(self.R, dtype=np.float64).reshape(4)))
        self._cost_bytes = self._cost_fingerprint()
        self._sig, self._lam = np.array([[.9, 0.0], [0.0, .9]]), .001
        self._sig_bytes = self._sig.tobytes()
        self._goal_bytes = None
        # fixed I/O buffers of the hot call (pointers resolved once)
        self._xin, self._gin, self._uout, self._xout = np.empty(3), np.empty(3), np.empty(2), np.empty(3)
        self._pxin, self._pgin, self._puout, self._pxout = (_capi.dptr(a) for a in (self._xin, self._gin, self._uout, self._xout))
        if self._noise_std_override is not None:
            _capi.check(self._lib.mppi_set_noise_std(self._h, _capi.dptr(self._noise_std_override)), "mppi_set_noise_std")

    def _cost_fingerprint(self):
        try:
            return self.Q.tobytes() + self.P1.tobytes() + self.R.tobytes()
        except AttributeError:       # somebody assigned a list: take the slow path
            return None

    def _sync_cost(self):
        """Q / R / P1 are public attributes in the reference; re-create the engine if they changed."""
        fp = self._cost_fingerprint()
        if fp is not None and fp == self._cost_bytes:
            return
        self._cost_bytes = fp
        key = (tuple(_diag3(self.Q, "Q")), tuple(_diag3(self.P1, "P1")),
               tuple(np.asarray(self.R, dtype=np.float64).reshape(4)))
        if key != self._cost_key:
            U = self.latest_uvec
            self.close()
            self._create()
            self.latest_uvec = U

    def _sync_sampling(self, sig, lam):
        if lam == self._lam and isinstance(sig, np.ndarray) and sig.dtype == np.float64 and sig.tobytes() == self._sig_bytes:
            return
        sig = _capi.f64(sig, (2, 2))
        self._sig_bytes = sig.tobytes()
        if lam != self._lam or not np.array_equal(sig, self._sig):
            _capi.check(self._lib.mppi_set_sampling(self._h, _capi.dptr(sig), float(lam)), "mppi_set_sampling")
            if self._noise_std_override is not None:
                _capi.check(self._lib.mppi_set_noise_std(self._h, _capi.dptr(self._noise_std_override)), "mppi_set_noise_std")
            self._sig, self._lam = sig.copy(), lam

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mppi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- latest_uvec lives on the device ------------------------------------------------------
    @property
    def latest_uvec(self):
        U = np.empty((2, self.horizon))
        _capi.check(self._lib.mppi_get_nominal(self._h, _capi.dptr(U)), "mppi_get_nominal")
        return U

    @latest_uvec.setter
    def latest_uvec(self, U):
        U = _capi.f64(U, (2, self.horizon))
        _capi.check(self._lib.mppi_set_nominal(self._h, _capi.dptr(U)), "mppi_set_nominal")

    # ---- reference methods ----------------------------------------------------------------------
    def initialize(self):
        """control/src/mppi:79-83."""
        self.fin_time = [0]
        _capi.check(self._lib.mppi_reset(self._h), "mppi_reset")          # latest_uvec = uvec_init (zeros)
        if np.any(self.uvec_init != 0):
            self.latest_uvec = copy.deepcopy(self.uvec_init)
        self.uvec = np.array([self.uvec_init[:, 0]])
        self.path = np.array([self.start])

    # `path` (n,3) and `uvec` (n,2) grow by one row per step (control/src/mppi:95-97); they are ndarray attributes in
    # the reference, here ndarray VIEWS of append-only logs (assigning an array re-seeds the log)
    @property
    def path(self):
        return self._path_log.array

    @path.setter
    def path(self, a):
        self._path_log = _Log.from_array(a)

    @property
    def uvec(self):
        return self._uvec_log.array

    @uvec.setter
    def uvec(self, a):
        self._uvec_log = _Log.from_array(a)

    def get_path(self, state, goal, sig=np.array([[.9, 0.0], [0.0, .9]]), lam=.001):
        """control/src/mppi:85-102 -- one MPPI step on the GPU; returns the predicted next state."""
        self._sync_cost()
        self._sync_sampling(sig, lam)
        self._xin[:] = state
        self._gin[:] = goal
        gb = self._gin.tobytes()
        if gb != self._goal_bytes:
            _capi.check(self._lib.mppi_set_goal(self._h, self._pgin), "mppi_set_goal")
            self._goal_bytes = gb
        st = self._lib.mppi_step(self._h, self._pxin, self._puout, self._pxout)
        if st != _capi.MPPI_OK:
            _capi.check(st, "mppi_step")
        x = self._xout.copy()
        self._path_log.append(x)                                                # :95
        self._uvec_log.append(self._uout)                                       # :96-97
        self.fin_time.append(self.fin_time[-1] + self.dt)                       # :98
        return x

    def solve_path(self, start, goal, sig=np.array([[1.0, 0.0], [0.0, 1.0]]), lam=.01, max_iters=None):
        """control/src/mppi:104-125 (max_iters is an added safety valve, default unlimited)."""
        self.start = start
        self.goal = goal
        state = start
        self.path = np.array([state])
        self.latest_uvec = copy.deepcopy(self.uvec_init)
        print("STARTING")
        i = 0
        tstart = time.time()
        while np.linalg.norm(state[:2] - goal[:2]) > self.thresh:
            i += 1
            state = self.get_path(state, goal, sig, lam)
            if i % 200 == 0:
                print("Iteration: {} \t State: {}".format(i, state))
            if max_iters is not None and i >= max_iters:
                break
        t_elapsed = time.time() - tstart
        print("Finished after {} iterations. Final State: {}, Time Taken: {}, Time Per Iter: {}".format(
            i, state, t_elapsed, t_elapsed / float(max(i, 1))))

    def get_cost2go(self, state, uvec, goal, lam, sig, eps=None):
        """control/src/mppi:127-178.  Returns (value_fcn (T,K), eps list of T arrays (2,K)).
        Like the reference, eps is drawn from the global NumPy stream (host side, this is the
        debug entry point) unless `eps` (T,2,K) is given (noise replay); the rollouts run on the device in fp64."""
        self._sync_cost()
        self._sync_sampling(sig, lam)
        T, K = self.horizon, self.samples
        if eps is None:
            eps = self.draw_noise()
        eps = _capi.f64(eps, (T, 2, K))
        V = np.empty((T, K))
        _capi.check(self._lib.mppi_cost_to_go(self._h, _capi.dptr(_capi.f64(state, (3,))), _capi.dptr(_capi.f64(uvec, (2, T))),
                                              _capi.dptr(_capi.f64(goal, (3,))), _capi.dptr(eps), _capi.dptr(V)),
                    "mppi_cost_to_go")
        return V, [eps[t] for t in range(T)]

    def update_action(self, uvec, eps, value_fcn, sig, lam):
        """control/src/mppi:186-208."""
        self._sync_sampling(sig, lam)
        T, K = self.horizon, self.samples
        out = np.empty((2, T))
        _capi.check(self._lib.mppi_update_action(self._h, _capi.dptr(_capi.f64(uvec, (2, T))),
                                                 _capi.dptr(_capi.f64(np.asarray(eps), (T, 2, K))),
                                                 _capi.dptr(_capi.f64(value_fcn, (T, K))), _capi.dptr(out)),
                    "mppi_update_action")
        return out

    def perform_action(self, state, uvec):
        """control/src/mppi:210-213."""
        out = np.empty(3)
        _capi.check(self._lib.mppi_perform_action(self._h, _capi.dptr(_capi.f64(state, (3,))),
                                                  _capi.dptr(_capi.f64(uvec, (2, self.horizon))), _capi.dptr(out)),
                    "mppi_perform_action")
        return out

    # ---- extensions (noise record/replay, grid, stats) -------------------------------------------
    def set_noise(self, eps):
        """Replay: use eps (T,2,K) for all following steps (SURVEY 8c direction i)."""
        eps = _capi.f64(np.asarray(eps), (self.horizon, 2, self.samples))
        _capi.check(self._lib.mppi_set_noise(self._h, _capi.dptr(eps)), "mppi_set_noise")

    def use_philox(self, seed=0):
        _capi.check(self._lib.mppi_use_philox(self._h, int(seed)), "mppi_use_philox")

    def get_noise(self):
        """Record: eps (T,2,K) the last step used (SURVEY 8c direction ii)."""
        eps = np.empty((self.horizon, 2, self.samples))
        _capi.check(self._lib.mppi_get_noise(self._h, _capi.dptr(eps)), "mppi_get_noise")
        return eps

    def draw_noise(self):
        """(T,2,K) noise drawn exactly as the reference draws it: T calls of
        np.random.normal(0, sig[0,0], size=(2,K)) on the global legacy stream (control/src/mppi:143-146)."""
        return np.stack([np.random.normal(0, self._sig[0, 0], size=(2, self.samples)) for _ in range(self.horizon)])

    def set_capture(self, on=True):
        _capi.check(self._lib.mppi_set_capture(self._h, 1 if on else 0), "mppi_set_capture")

    def get_value_fcn(self):
        V = np.empty((self.horizon, self.samples))
        _capi.check(self._lib.mppi_get_cost_to_go(self._h, _capi.dptr(V)), "mppi_get_cost_to_go")
        return V

    def get_last_update(self):
        U = np.empty((2, self.horizon))
        _capi.check(self._lib.mppi_get_last_update(self._h, _capi.dptr(U)), "mppi_get_last_update")
        return U

    def set_grid(self, cells, res, origin, w_obs):
        """NEW: int8 occupancy grid (H,W) row-major with the map package's conventions
        (map/src/map/grid.cpp:126-144,251-266; nav_msgs/OccupancyGrid as published at map/src/viz_grid.cpp:109-137)."""
        cells = np.ascontiguousarray(np.asarray(cells, dtype=np.int8))
        H, W = cells.shape
        _capi.check(self._lib.mppi_set_grid(self._h, cells.ctypes.data_as(C.POINTER(C.c_int8)), W, H, float(res),
                                            float(origin[0]), float(origin[1]), float(w_obs)), "mppi_set_grid")

    def update_grid(self, patch, x0, y0):
        """NEW: overwrite the rectangle [y0:y0+h, x0:x0+w] of the resident grid with `patch` (h, w) int8 -- the incremental map
        updates of map/src/map/grid.cpp:155-199 -- asynchronously, without re-creating anything."""
        patch = np.ascontiguousarray(np.asarray(patch, dtype=np.int8))
        h, w = patch.shape
        _capi.check(self._lib.mppi_update_grid(self._h, patch.ctypes.data_as(C.POINTER(C.c_int8)), int(x0), int(y0), w, h),
                    "mppi_update_grid")

    def clear_grid(self):
        _capi.check(self._lib.mppi_clear_grid(self._h), "mppi_clear_grid")

    def stats(self):
        t = _capi.MppiTiming()
        _capi.check(self._lib.mppi_last_stats(self._h, C.byref(t)), "mppi_last_stats")
        return {k: getattr(t, k) for k, _ in t._fields_}

    def io_bytes(self):
        a, b = C.c_size_t(), C.c_size_t()
        _capi.check(self._lib.mppi_io_bytes(self._h, C.byref(a), C.byref(b)), "mppi_io_bytes")
        return a.value, b.value

    def launch_info(self):
        info = (C.c_int32 * 8)()
        _capi.check(self._lib.mppi_launch_info(self._h, info), "mppi_launch_info")
        d = dict(zip(["block", "grid", "tiles", "smem_bytes", "ctas_per_sm", "regs"], list(info)[:6]))
        d["variant"] = ("general", "fast", "lean")[info[6]]
        d["records_per_step"] = info[7]     # partial records per time step (== grid, or one per 64-rollout tile of the SM-wide kernel)
        return d

    def bench(self, x0, steps=20, warmup=3, flush_l2=True, per_kernel=True):
        t = _capi.MppiTiming()
        _capi.check(self._lib.mppi_set_goal(self._h, _capi.dptr(_capi.f64(self.goal, (3,)))), "mppi_set_goal")
        self._goal_bytes = None   # the engine's goal no longer matches what get_path last sent
        _capi.check(self._lib.mppi_bench(self._h, _capi.dptr(_capi.f64(x0, (3,))), steps, warmup, int(flush_l2),
                                         int(per_kernel), C.byref(t)), "mppi_bench")
        return {k: getattr(t, k) for k, _ in t._fields_}
