"""Compare experimental kernel builds (motion_planning_b200/lib/libmppi_b200_<tag>.so) on one config.
   python profiles/variants.py [K] [T] [tags...]"""
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
K = sys.argv[1] if len(sys.argv) > 1 else "65536"
T = sys.argv[2] if len(sys.argv) > 2 else "64"
tags = sys.argv[3:] or ["", "block64", "p7"]
code = r'''
import sys, os, json
sys.path.insert(0, %r)
import numpy as np, motion_planning_b200 as mp
K, T = int(sys.argv[1]), int(sys.argv[2])
out = {}
for prec in ("mixed", "f32"):
    m = mp.MPPI(horizon=T, samples=K, precision=prec, seed=0); m.goal = np.array([0., -1., 0.])
    r = m.bench(np.zeros(3), steps=40, warmup=5, flush_l2=True, per_kernel=True)
    out[prec] = (round(r["step_ms"] * 1e3, 2), round(r["rollout_ms"] * 1e3, 2), m.launch_info()["block"], m.launch_info()["regs"], m.launch_info()["variant"],
                 r["refine_candidates"], float("%%.2g" %% r["refine_max_dev"]))
    m.close()
print(json.dumps(out))
''' % ROOT
for tag in tags:
    env = dict(os.environ)
    name = tag
    if tag.startswith("block"):
        env["MPPI_B200_BLOCK"] = tag[5:]
    elif tag in ("fast", "general"):
        env["MPPI_B200_VARIANT"] = tag
    elif tag:
        env["MPPI_B200_LIB"] = os.path.join(ROOT, "motion_planning_b200", "lib", "libmppi_b200_%s.so" % tag)
    r = subprocess.run([sys.executable, "-c", code, K, T], env=env, capture_output=True, text=True)
    print("%-8s K=%s T=%s (step us, rollout us, block, regs, variant, candidates, max|V32-V64|): %s %s" % (name or "default", K, T, r.stdout.strip(), r.stderr.strip()[-300:]))
