"""What a caller-supplied KINEMATIC functor costs at the headline size (K=65536, T=64, precision mixed), next to the built-in
diff-drive on the same (general) code path and on the product (lean, SM-wide) path -- and whether the NVRTC instantiation
reproduces the nvcc-built one bit for bit.      python profiles/kinematic_timing.py > profiles/r02_kinematic_timing.txt"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import user_models as um                     # noqa: E402
import motion_planning_b200 as M             # noqa: E402

K, T = 65536, 64
PARK = np.array([0.0, -1.0, 0.0])
kin = M.KinematicModel(um.DD_KIN_CUDA, name="dd_kin", **um.DD_KIN_BOUNDS)
ode = M.UserModel(um.DD_CUDA, name="dd_ode")
os.environ["MPPI_B200_VARIANT"] = "general"
built_general = M.MPPI(horizon=T, samples=K, precision="mixed", seed=1)
del os.environ["MPPI_B200_VARIANT"]
engines = [("built-in diff-drive, product path (mixed)", M.MPPI(horizon=T, samples=K, precision="mixed", seed=1)),
           ("built-in diff-drive, general path (mixed)", built_general),
           ("kinematic functor via NVRTC (mixed)", M.MPPI(model=kin, horizon=T, samples=K, precision="mixed", seed=1)),
           ("kinematic functor via NVRTC (f64)", M.MPPI(model=kin, horizon=T, samples=K, precision="f64", seed=1)),
           ("ODE functor via NVRTC (f64)", M.MPPI(model=ode, horizon=T, samples=K, precision="f64", seed=1)),
           ("ODE functor via NVRTC (f32)", M.MPPI(model=ode, horizon=T, samples=K, precision="f32", seed=1))]
s0 = np.array([0.1, 0.0, 0.4])
U = {}
for name, m in engines:
    m.get_path(s0, PARK)
    U[name] = m.latest_uvec.copy()
a, b = U["kinematic functor via NVRTC (mixed)"], U["built-in diff-drive, general path (mixed)"]
print("first step, kinematic functor vs built-in on the general path: bit-identical %s, max |dU| %.3g" % (np.array_equal(a, b), np.max(np.abs(a - b))))
print("first step, kinematic functor vs built-in on the product path: max |dU| %.3g" % np.max(np.abs(a - U["built-in diff-drive, product path (mixed)"])))
print("K=%d T=%d, device-resident closed loop, L2 flushed between steps, us per step (rollout | reduce incl. finalize):" % (K, T))
for name, m in engines:
    m.initialize()
    m.goal = PARK
    r = m.bench(s0, steps=30, warmup=5, flush_l2=True, per_kernel=True)
    print("  %-44s %7.2f   (%6.2f | %6.2f)   %s" % (name, 1e3 * r["step_ms"], 1e3 * r["rollout_ms"], 1e3 * (r["reduce_ms"] + r["finalize_ms"]), m.launch_info()))
    m.close()
