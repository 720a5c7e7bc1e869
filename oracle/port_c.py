"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/_build/libmppi_port.so (the C/OpenMP port)."""
import ctypes as C
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libmppi_port.so")
_lib = None
_dp = C.POINTER(C.c_double)


class PortParams(C.Structure):
    _fields_ = [("K", C.c_int), ("T", C.c_int), ("dt", C.c_double), ("q", C.c_double * 3), ("R", C.c_double * 4),
                ("p1", C.c_double * 3), ("sig", C.c_double * 4), ("lam", C.c_double), ("u_max", C.c_double * 2),
                ("r", C.c_double), ("L", C.c_double), ("eps_floor", C.c_double), ("noise_std", C.c_double)]


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        assert _lib.port_sizeof_params() == C.sizeof(PortParams)
        _lib.port_sizeof_rk_state.restype = C.c_size_t
        _lib.port_gauss.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


class Port(object):
    def __init__(self, K, T, seed=0):
        self.l = lib()
        self.p = PortParams()
        self.l.port_default_params(C.byref(self.p), K, T)
        self.K, self.T = K, T
        self.rs = C.create_string_buffer(self.l.port_sizeof_rk_state())
        self.l.port_seed(self.rs, C.c_uint32(seed))
        self.U = np.zeros((2, T))
        self.w_eps = np.empty((T, 2, K))
        self.w_V = np.empty((T, K))

    def normal(self, scale, n):
        out = np.empty(n)
        self.l.port_normal(self.rs, C.c_double(scale), _p(out), C.c_long(n))
        return out

    def step(self, x0, goal, noise_mode=1, eps=None, seed=0):
        """returns dict(u0, x_next, U_new, U_shift, V (as left by update_action), eps)"""
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        goal = np.ascontiguousarray(goal, dtype=np.float64)
        u0, xn, Un = np.empty(2), np.empty(3), np.empty((2, self.T))
        e = None if eps is None else np.ascontiguousarray(eps, dtype=np.float64)
        self.l.port_step(C.byref(self.p), _p(x0), _p(goal), _p(self.U), C.c_int(noise_mode),
                         _p(e) if e is not None else None, self.rs, C.c_uint64(seed), _p(self.w_eps), _p(self.w_V),
                         _p(u0), _p(xn), _p(Un))
        return dict(u0=u0, x_next=xn, U_new=Un, U_shift=self.U.copy(), eps=self.w_eps if e is None else e)

    def cost2go(self, x0, U, goal, eps):
        V = np.empty((self.T, self.K))
        self.l.port_cost2go(C.byref(self.p), _p(np.ascontiguousarray(x0, dtype=np.float64)),
                            _p(np.ascontiguousarray(U, dtype=np.float64)), _p(np.ascontiguousarray(goal, dtype=np.float64)),
                            _p(np.ascontiguousarray(eps, dtype=np.float64)), _p(V))
        return V


def time_workload(K, T, budget_s=12.0):
    """cpu_baseline leg of bench.py: whole get_path steps of the C/OpenMP port on all host cores."""
    port = Port(K, T)
    s = np.zeros(3)
    goal = np.array([0.0, -1.0, 0.0])
    out = port.step(s, goal, noise_mode=2, seed=1)       # warm-up (page faults of the 100 MB scratch)
    s = out["x_next"]
    t0, n = time.perf_counter(), 0
    while True:
        out = port.step(s, goal, noise_mode=2, seed=n + 2)
        s = out["x_next"]
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 200:
            break
    dt = (time.perf_counter() - t0) / n
    nt = lib().port_num_threads()
    return {"value": K / dt, "unit": "rollouts/s", "cores": nt, "kind": "port",
            "sample": "C/OpenMP f64 restatement oracle/mppi_port.c (same op order as control/src/mppi, K loop threaded, "
                      "parallel counter-based noise inside the timed region), full K=%d T=%d, %d steps, %.4f s/step on %d threads"
                      % (K, T, n, dt, nt)}
