"""Generate tests/golden/ref_*.npz from the UNMODIFIED reference (control/src/mppi).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference ships no golden vectors of its own, so these
outputs of the live reference are what pins the oracle and the CUDA path.

Each case: fresh module load (=> np.random.seed(0), control/src/mppi:15), fresh
MPPI(horizon=T, samples=K), closed loop on the model: s = m.get_path(s, goal).
Stored per iteration: u0 = m.uvec[-1], x_next = returned state, U = m.latest_uvec
(after the shift), and for the first iteration the noise eps (T,2,K) and the
cost-to-go V (T,K) captured by wrapping get_cost2go (wrapper only records).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_controller, ref_loader  # noqa: E402

CASES = [
    # name,            K,    T,   x0,              goal,             iters, store_big
    ("c1_park_k128_t32", 128, 32, (0.0, 0.0, 0.0), (0.0, -1.0, 0.0), 4, True),
    ("node_default_k10_t100", 10, 100, (0.0, 0.0, 0.0), (0.0, -1.0, 0.0), 3, True),
    ("pentagon_leg0_k128_t32", 128, 32, (0.0, 0.0, 0.0), (1.0, 0.0, 0.0), 3, True),
    ("park_k1024_t64", 1024, 64, (0.0, 0.0, 0.0), (0.0, -1.0, 0.0), 2, False),
    ("offset_start_k256_t16", 256, 16, (0.3, -0.2, 2.9), (1.0, 0.5, -1.0), 3, True),
]


def run_case(name, K, T, x0, goal, iters, store_big):
    ref = ref_loader.load_reference()          # re-seeds the legacy global stream with 0
    m = ref.MPPI(horizon=T, samples=K)
    captured = {}
    orig = m.get_cost2go

    def recording(state, uvec, goal_, lam, sig):
        V, eps = orig(state, uvec, goal_, lam, sig)
        if "V" not in captured:
            captured["V"] = np.array(V, dtype=np.float64).copy()
            captured["eps"] = np.array(eps, dtype=np.float64).copy()
        return V, eps

    m.get_cost2go = recording
    s = np.array(x0, dtype=np.float64)
    g = np.array(goal, dtype=np.float64)
    u0s, xs, Us = [], [], []
    for _ in range(iters):
        s = m.get_path(s, g)
        u0s.append(np.array(m.uvec[-1]))
        xs.append(np.array(s))
        Us.append(np.array(m.latest_uvec))
    out = dict(K=K, T=T, x0=np.array(x0), goal=g, u0=np.array(u0s), x_next=np.array(xs), U_shift=np.array(Us),
               numpy_version=np.__version__)
    if store_big:
        out["V0"] = captured["V"]
        out["eps0"] = captured["eps"]
    else:
        out["V0_row0_min"] = captured["V"][0].min()
        out["V0_row0_max"] = captured["V"][0].max()
    np.savez_compressed(os.path.join(HERE, "ref_%s.npz" % name), **out)
    print(name, "u0[0] =", repr(u0s[0]))


# the caller of the hot path: the reference ROS node (control/src/mppi:296-389) driven without ROS (oracle/ref_controller.py)
CONTROLLER_CASES = [
    # name,                   waypoints ("waypoints" ROS parameter; None = parallel park), K,  T, pose0,           callbacks
    ("waypoints_k32_t16", [[0.1, 0.0], [0.1, 0.1], [0.0, 0.0]], 32, 16, (0.0, 0.0, 0.0), 100),
    ("park_k32_t16", None, 32, 16, (0.0, -0.93, 0.4), 60),
]


def unicycle_plant(pose0, dt):
    """Constant-twist arc over dt (the simulated robot of control/launch/mppi_pentagon.launch; generator-side copy so the
    fixture does not depend on product code)."""
    import math
    st = {"p": np.array(pose0, dtype=np.float64)}

    def step(vx, wz):
        x, y, th = st["p"]
        a = wz * dt
        if abs(a) < 1e-12:
            x += vx * dt * math.cos(th)
            y += vx * dt * math.sin(th)
        else:
            r = vx / wz
            x += r * (math.sin(th + a) - math.sin(th))
            y -= r * (math.cos(th + a) - math.cos(th))
        th = math.atan2(math.sin(th + a), math.cos(th + a))
        st["p"] = np.array([x, y, th])
        return st["p"]
    return step


def run_controller_case(name, waypoints, K, T, pose0, n):
    ref, node, log = ref_controller.load_node(waypoints, dict(horizon=T, samples=K))
    # the noise is NOT stored: it is the legacy global NumPy stream seeded with 0 at module load (control/src/mppi:15),
    # T draws of shape (2,K) per get_path (:143-146) -- the tests re-draw it (MPPI.draw_noise) callback by callback
    poses, twists, flags, nominal = ref_controller.run_node(node, unicycle_plant(pose0, 1.0 / T), pose0, n)
    np.savez_compressed(os.path.join(HERE, "node_%s.npz" % name), K=K, T=T, waypoints=np.array(waypoints if waypoints else []),
                        poses=poses, twists=twists, flags=flags, nominal=nominal, uvec_rows=node.mppi.uvec.shape[0], n_log=len(log), numpy_version=np.__version__)
    print(name, "final flags", flags[-1], "any done", int(flags[:, 2].max()), "max idx", int(flags[:, 0].max()), "last twist", twists[-1])


def rng_kat():
    np.random.seed(0)
    a = np.random.normal(size=4)
    b = np.random.normal(0, .9, size=(2, 5))
    np.savez(os.path.join(HERE, "ref_rng_kat.npz"), normal4=a, normal_2x5=b)


if __name__ == "__main__":
    if ref_loader.available() != "source":
        sys.exit("needs /root/reference (build container only)")
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    if only in ("", "mppi"):
        for c in CASES:
            run_case(*c)
        rng_kat()
    if only in ("", "controller"):
        for c in CONTROLLER_CASES:
            run_controller_case(*c)
