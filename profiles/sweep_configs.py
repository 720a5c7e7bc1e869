"""All BASELINE.json configs on ONE GPU (device-resident loop): timing + refinement statistics.
   python profiles/sweep_configs.py [steps]   -> one JSON line per config"""
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402
import motion_planning_b200 as mp  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30


def demo_grid():
    """the map package's demo grid (map/config/map.yaml at scale 5, resolution 0.06, inflate 0.1: map/launch/viz_map.launch:52-57)
    as produced by the reference's own Grid::build_map (tests/golden/map_grid_reference.npz, tests/golden/make_map_golden.py)"""
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "map_grid_reference.npz"))
    return g["demo"], float(g["demo_res"]), g["demo_origin"]


CONFIGS = [
    ("c1 diff-drive park K=128 T=32", dict(horizon=32, samples=128), (0, 0, 0), (0, -1, 0), None),
    ("c2 diff-drive park K=65536 T=64", dict(horizon=64, samples=65536), (0, 0, 0), (0, -1, 0), None),
    ("c3 bicycle pentagon leg K=65536 T=64", dict(horizon=64, samples=65536, model=mp.bicycle_rk4, u_max=[0.22, 0.6],
                                                  noise_std=[0.08, 0.25]), (0, 0, 0), (1, 0, 0), None),
    ("c3b diff-drive pentagon leg K=65536 T=64", dict(horizon=64, samples=65536), (0, 0, 0), (1, 0, 0), None),
    ("c4 diff-drive + map-package grid K=262144 T=64", dict(horizon=64, samples=262144), (1.0, 1.5, np.pi), (1.0, 0.5, -np.pi / 2), "grid"),
    ("c5/8 diff-drive park K=262144 T=128 (one GPU's share of config 5)", dict(horizon=128, samples=262144), (0, 0, 0), (0, -1, 0), None),
    ("c5 diff-drive park K=2097152 T=128 on ONE GPU", dict(horizon=128, samples=2097152), (0, 0, 0), (0, -1, 0), None),
]
for name, kw, x0, goal, extra in CONFIGS:
    for prec in ("mixed", "f32", "f64"):
        if prec == "f64" and kw["samples"] > 300000:
            continue
        m = mp.MPPI(precision=prec, seed=0, **kw)
        m.goal = np.array(goal, dtype=np.float64)
        if extra == "grid":
            cells, res, origin = demo_grid()
            m.set_grid(cells, res, origin, 250.0)
            m.latest_uvec = np.full((2, kw["horizon"]), 5.0)       # driving towards obstacle D: the grid term is in play
        r = m.bench(np.array(x0, dtype=np.float64), steps=steps, warmup=3, flush_l2=True, per_kernel=True)
        K, T = kw["samples"], kw["horizon"]
        print(json.dumps({"config": name, "precision": prec, "K": K, "T": T, "ms_per_step": r["step_ms"],
                          "rollouts_per_s": K / (r["step_ms"] * 1e-3), "state_steps_per_s": K * T / (r["step_ms"] * 1e-3),
                          "rollout_ms": r["rollout_ms"], "reduce_finalize_ms": r["reduce_ms"],
                          "refine_candidates": r["refine_candidates"], "refine_overflow_steps": r["refine_overflow"],
                          "refine_max_dev": r["refine_max_dev"], "refine_head_room": r["refine_head_room"], "launch": m.launch_info()}))
        m.close()
