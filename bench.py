#!/usr/bin/env python
"""bench.py -- MPPI rollouts/s on B200 (BASELINE.json metric) with roofline, e2e and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one complete MPPI.get_path (control/src/mppi:85-102): noise -> K rollouts x T RK4 steps ->
costs -> per-t soft-min weights -> control update -> clip / Savitzky-Golay / clip -> apply & shift.

Workload (N=1): BASELINE.json configs[1] = diff-drive parallel-park, K=65536, T=64, 1 x B200.  For
N > 1 the rollouts are sharded over ranks with per-GPU K fixed (weak scaling, K_total = N * 65536)
and one 3 KB record exchanged per step; that line's `value` is comparable with N=1.  EVERY line
(N = 1 too) additionally carries a `config5` block: BASELINE.json configs[4], K_total = 2097152,
T = 128 sharded over the N GPUs of the run (strong scaling), and for N > 1 a `parity` record: the
sharded engines against ONE engine at K_total on rank 0, same noise, two closed-loop steps.

value   = rollouts/s with the controller state resident in HBM (closed loop on the model entirely on
          the device), CUDA events on the launch stream, L2 flushed between timed steps.
e2e     = the same metric through the public API call MPPI.get_path / mppi_step with HOST x0 in and
          (u_t, x_next) out every step, copies inside the timed region.
roofline= the fused rollout+cost kernel against the fp32 FMA peak measured in the same run
          (this path is fp32-ALU/SFU bound, not HBM bound: SURVEY.md 8d) + the HBM view for contrast.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_PER_GPU, T_HORIZON = 65536, 64
K5_TOTAL, T5 = 2097152, 128                # BASELINE.json configs[4]: sharded over the GPUs of the run
GOAL = np.array([0.0, -1.0, 0.0])          # parallel park, control/src/mppi:337
X0 = np.array([0.0, 0.0, 0.0])
F_ALG = 133.0                              # algorithmic FLOP per (rollout, step), SURVEY.md 8(d)
METRIC = "MPPI rollouts/sec (K x T states/sec = value * T)"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ CPU legs
def cpu_baseline_port(budget_s=12.0):
    """The oracle port timed on the host cores (checker only, never the product path): the C/OpenMP
    restatement when it is built (all cores), else the vectorised NumPy restatement (1 core)."""
    try:
        from oracle import port_c
        if port_c.available():
            return port_c.time_workload(K_PER_GPU, T_HORIZON, budget_s)
    except ImportError:
        pass
    from oracle import mppi_oracle as orc
    Ks = 16384
    p = orc.Params(K=Ks, T=T_HORIZON)
    rng = np.random.RandomState(0)
    U, s = np.zeros((2, T_HORIZON)), X0.copy()
    t0, n = time.perf_counter(), 0
    while True:
        eps = rng.normal(0, 0.9, size=(T_HORIZON, 2, Ks))
        out = orc.step(p, s, GOAL, U, eps)
        s, U = out["x_next"], out["U_shift"]
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 50:
            break
    dt = (time.perf_counter() - t0) / n
    return {"value": Ks / dt, "unit": "rollouts/s", "cores": 1, "kind": "port",
            "sample": "NumPy f64 restatement (oracle/mppi_oracle.py, noise drawn by NumPy inside the timed region), "
                      "%d of the %d rollouts x T=%d, %d steps, %.3f s/step" % (Ks, K_PER_GPU, T_HORIZON, n, dt)}


def run_reference_arm(args):
    """--impl reference: the UNMODIFIED reference MPPI class (control/src/mppi, byte-compiled into
    oracle/_ref/mppi.pyc) on the host CPU.  It is single-threaded Python by construction (1 core).
    Each step is a bounded sample of the workload: Ks of the K rollouts at the full horizon T."""
    from oracle import ref_loader
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    kind, cores = "reference", 1
    # ~0.6 ms per rollout at T=64 (8.7 us per (k,t)): keep the whole run near 60-90 s
    Ks = int(max(64, min(2048, (75.0 / max(steps + warmup, 1)) / 0.6e-3)))
    Ks -= Ks % 2
    if ref_loader.available():
        ref = ref_loader.load_reference()
        m = ref.MPPI(horizon=T_HORIZON, samples=Ks)
        stepper = lambda s: m.get_path(s, GOAL)          # noqa: E731
        what = "unmodified reference control/src/mppi (%s)" % ref.__ref_kind__
    else:
        from oracle import mppi_oracle as orc
        kind = "port"
        p = orc.Params(K=Ks, T=T_HORIZON)
        st = {"U": np.zeros((2, T_HORIZON))}

        def stepper(s):
            out = orc.step(p, s, GOAL, st["U"], orc.draw_reference_noise(p))
            st["U"] = out["U_shift"]
            return out["x_next"]
        what = "NumPy restatement oracle/mppi_oracle.py (reference not loadable here)"
    s = X0.copy()
    for _ in range(warmup):
        s = stepper(s)
    t0 = time.perf_counter()
    for _ in range(steps):
        s = stepper(s)
    dt = (time.perf_counter() - t0) / steps
    val = Ks / dt
    # the reference's get_path is a Python loop over the K samples (control/src/mppi:158-161): linear in K.  One step at the
    # FULL K of the workload is timed as well, so the sampled figure is checked against the real configuration in this run
    # (--ref-full times every step at full K instead; ~15-40 s per step)
    full = None
    if ref_loader.available() and (args.ref_full or args.ref_full_step):
        mf = ref.MPPI(horizon=T_HORIZON, samples=K_PER_GPU)
        nfull = steps if args.ref_full else 1
        tf0 = time.perf_counter()
        sf = X0.copy()
        for _ in range(nfull):
            sf = mf.get_path(sf, GOAL)
        tfull = (time.perf_counter() - tf0) / nfull
        full = {"K": K_PER_GPU, "T": T_HORIZON, "steps": nfull, "s_per_step": tfull, "rollouts_per_s": K_PER_GPU / tfull,
                "sampled_over_full": val / (K_PER_GPU / tfull)}
        if args.ref_full:
            val, dt, Ks = K_PER_GPU / tfull, tfull, K_PER_GPU
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "rollouts/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "diff-drive parallel-park K=%d T=%d (BASELINE.json configs[1])" % (K_PER_GPU, T_HORIZON),
                   "K": K_PER_GPU, "T": T_HORIZON, "sample_K": Ks, "full_K_step": full,
                   "note": "value = rollouts/s of get_path; every timed step rolls sample_K of the K samples at the full horizon "
                           "(the reference is linear in K); full_K_step = one get_path at the full K timed in this run"},
        "cpu_baseline": {"value": val, "unit": "rollouts/s", "cores": cores, "kind": kind,
                         "sample": "%s; %d of the %d rollouts per step at T=%d, %d timed steps" % (what, Ks, K_PER_GPU, T_HORIZON, steps)},
        "e2e": {"value": val, "unit": "rollouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ ours
def run_single(args):
    import motion_planning_b200 as mp
    from motion_planning_b200 import _capi
    import ctypes as C
    lib = _capi.load()
    if lib.mppi_device_count() < 1:
        sys.exit("bench.py: no CUDA device and no CPU fallback (use --impl reference for the CPU arm)")
    K, T = K_PER_GPU, T_HORIZON
    tf, mhz = C.c_double(), C.c_double()
    _capi.check(lib.mppi_measure_fp32_peak(0, C.byref(tf), C.byref(mhz)), "mppi_measure_fp32_peak")
    results = {}
    sampler = ClockSampler(0)                            # clocks / throttle reasons DURING all timed regions
    for prec in ("f32", "f64", args.precision):          # headline precision last
        m = mp.MPPI(horizon=T, samples=K, precision=prec, seed=0)
        m.goal = GOAL
        r = m.bench(X0, steps=args.steps, warmup=args.warmup, flush_l2=True, per_kernel=True)
        # the same device-resident loop WITHOUT the flush: what a controller stepping back to back sees (instructions and the
        # few KB of state stay in L2); reported next to the flushed figure, never instead of it
        r["warm_step_ms"] = m.bench(X0, steps=args.steps, warmup=args.warmup, flush_l2=False, per_kernel=False)["step_ms"]
        # e2e: the public call with HOST buffers, copies inside the timed region.
        # (1) through the Python mirror of the reference class (MPPI.get_path, what Controller calls)
        # Every timed call starts from a cold L2 and an idle GPU (mppi_debug_flush_l2 outside the timed intervals).
        flush = lambda: _capi.check(lib.mppi_debug_flush_l2(m._h), "mppi_debug_flush_l2")  # noqa: E731
        m.initialize()
        s = X0.copy()
        for _ in range(args.warmup):
            s = m.get_path(s, GOAL)
        acc = 0.0
        for _ in range(args.steps):
            flush()
            t0 = time.perf_counter()
            s = m.get_path(s, GOAL)
            acc += time.perf_counter() - t0
        r["e2e_shim_ms"] = acc / args.steps * 1e3
        # (2) through the C ABI directly (mppi_step with host double[3] in, double[2]+double[3] out): what a
        #     C/C++ host pays; same copies, no interpreter overhead around the call
        m.initialize()
        x, u, xn = X0.copy(), np.empty(2), np.empty(3)
        px, pu, pn = _capi.dptr(x), _capi.dptr(u), _capi.dptr(xn)
        step, h = lib.mppi_step, m._h
        _capi.check(lib.mppi_set_goal(h, _capi.dptr(GOAL)), "mppi_set_goal")
        for _ in range(args.warmup):
            step(h, px, pu, pn)
            x[:] = xn
        acc = 0.0
        for _ in range(args.steps):
            flush()
            t0 = time.perf_counter()
            if step(h, px, pu, pn) != 0:
                raise RuntimeError("mppi_step failed")
            acc += time.perf_counter() - t0
            x[:] = xn
        r["e2e_ms"] = acc / args.steps * 1e3
        t0 = time.perf_counter()               # the same loop back to back (warm L2), for reference
        for _ in range(args.steps):
            step(h, px, pu, pn)
            x[:] = xn
        r["e2e_warm_ms"] = (time.perf_counter() - t0) / args.steps * 1e3
        r["io"] = m.io_bytes()
        r["launch"] = m.launch_info()
        r["stats"] = m.stats()
        results[prec] = r
        m.close()
    # BASELINE.json configs[4] on this one GPU (the N=1 point of its strong-scaling curve; also where the fixed costs of the
    # rollout kernel vanish and its roofline fraction is best read)
    c5 = None
    if not args.no_config5:
        m5 = mp.MPPI(horizon=T5, samples=K5_TOTAL, precision=args.precision, seed=0)
        m5.goal = GOAL
        r5 = m5.bench(X0, steps=max(3, min(args.steps, 30)), warmup=3, flush_l2=True, per_kernel=True)
        ach5 = F_ALG * K5_TOTAL * T5 / (r5["rollout_ms"] * 1e-3) / 1e12
        c5 = {"workload": "diff-drive parallel-park K_total=%d T=%d (BASELINE.json configs[4]) on %d GPU" % (K5_TOTAL, T5, 1),
              "n_gpus": 1, "scaling": "strong", "ms_per_step": r5["step_ms"], "value": K5_TOTAL / (r5["step_ms"] * 1e-3),
              "unit": "rollouts/s", "state_steps_per_s": K5_TOTAL * T5 / (r5["step_ms"] * 1e-3), "steps": r5["steps"],
              "rollout_ms": r5["rollout_ms"], "roofline_frac_rollout_kernel": ach5 / tf.value if tf.value else None,
              "achieved_TFLOPs": ach5, "launch": m5.launch_info(),
              "refine": {"candidates_last_step": r5["refine_candidates"], "overflow_steps": r5["refine_overflow"],
                         "max_abs_dev_fp32_vs_fp64": r5["refine_max_dev"], "head_room": r5["refine_head_room"]}}
        m5.close()
    clocks = sampler.stop()
    r = results[args.precision]
    ms = r["step_ms"]
    value = K / (ms * 1e-3)
    achieved = F_ALG * K * T / (r["rollout_ms"] * 1e-3) / 1e12
    nrec = r["launch"].get("records_per_step") or r["launch"]["grid"]
    hbm_alg_bytes = 16 * T * nrec * 3 + 16 * T     # per-tile partial records (meta + floor sums + candidates) out, nominal in
    kname = "rollout_lean_sm_kernel" if (r["launch"].get("variant") == "lean" and r["launch"]["block"] == 512) else \
        "rollout_%s_kernel" % r["launch"].get("variant", "?")
    cpu = cpu_baseline_port() if not args.no_cpu else None
    traffic, traffic_file = None, None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
        cands = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_ncu_summary.json"))
        traffic_file = "profiles/" + cands[-1]
        ncu = json.load(open(os.path.join(ROOT, traffic_file)))
        key = [k for k in ncu if k.startswith("rollout") and ("precision %s" % args.precision) in k and ("K %d T %d" % (K, T)) in k]
        if key:
            m_ = ncu[key[0]]
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            traffic = sum(m_[n]["value"] * mult.get(m_[n]["unit"], 1) for n in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except Exception:
        traffic = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "rollouts/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"mixed": "f32 rollouts + f64 re-evaluation of the softmin support", "f32": "f32", "f64": "f64"}[args.precision],
        "data": "synthetic",
        "config": {"workload": "diff-drive parallel-park K=%d T=%d (BASELINE.json configs[1])" % (K, T), "K": K, "T": T,
                   "precision": args.precision, "weighting": "cost_to_go (reference)", "noise": "Philox4x32-10 in registers, 6 normals (3 steps x 2 channels) per call",
                   "l2": "flushed between timed steps, outside the timed intervals: 256 MiB overwritten, then 256 MiB of clean lines read (the step starts cold but is not charged the write-back of the flush's own dirty lines)",
                   "loop": "closed loop on the model, state resident in HBM", "launch": r["launch"]},
        "state_steps_per_s": value * T,
        "warm_l2": {"ms_per_step": r["warm_step_ms"], "value": K / (r["warm_step_ms"] * 1e-3),
                    "note": "same device-resident loop, L2 NOT flushed between steps (back-to-back operation): the flushed figure above "
                            "charges every step the refetch of the kernels' instructions and of the controller state from HBM"},
        "e2e": {"value": K / (r["e2e_ms"] * 1e-3), "unit": "rollouts/s", "h2d_bytes_per_step": r["io"][0],
                "d2h_bytes_per_step": r["io"][1], "ms_per_step": r["e2e_ms"],
                "call": "mppi_step (C ABI) with host x0 in / (u, x_next) out, closed loop on the host; x0 and goal ride in the "
                        "kernel arguments, the result block is stored by the finalize phase into mapped pinned host memory",
                "l2": "flushed before every timed call (outside the timed interval)", "warm_l2_ms_per_step": r["e2e_warm_ms"],
                "python_shim_ms_per_step": r["e2e_shim_ms"], "python_shim_value": K / (r["e2e_shim_ms"] * 1e-3)},
        "gpu_launches": r["launches"],
        # two launches per step; the finalize phase runs in the reduce kernel's finalizer block (eager launches with events in
        # between, so the PDL overlap of the two kernels is NOT in these two figures: their sum exceeds ms_per_step)
        "kernels_ms": {"rollout": r["rollout_ms"], "reduce_incl_finalize": r["reduce_ms"]},
        "roofline": {"bound": "fp32_alu", "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s",
                     "frac": achieved / tf.value if tf.value else None, "traffic": traffic,
                     "traffic_source": "from profile, NOT measured in this run: %s (dram__bytes_read.sum + dram__bytes_write.sum of one "
                                       "ncu --set full capture of this kernel at this configuration, caches flushed by ncu before the launch); "
                                       "most of the ~3 MB of per-tile partial records stay in L2" % traffic_file,
                     "kernel": kname, "flop_per_state_step": F_ALG,
                     "peak_source": "FFMA chain measured in this run (mppi_measure_fp32_peak); MEASURED_PEAKS.json has no fp32 entry",
                     "hbm_view": {"algorithmic_bytes": hbm_alg_bytes,
                                  "achieved_GBps": hbm_alg_bytes / (r["rollout_ms"] * 1e-3) / 1e9,
                                  "peak_GBps": peaks.get("hbm_gbs"), "note": "not HBM bound: nothing of size K*T leaves the SM"}},
        "refine": {"candidates_last_step": r["refine_candidates"], "overflow_steps": r["refine_overflow"],
                   "max_abs_dev_fp32_vs_fp64": r["refine_max_dev"], "head_room": r["refine_head_room"]},
        "config5": c5,
        "other_precisions": {p: {"value": K / (v["step_ms"] * 1e-3), "ms_per_step": v["step_ms"], "rollout_ms": v["rollout_ms"],
                                 "e2e_ms": v["e2e_ms"]} for p, v in results.items() if p != args.precision},
        "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))


def _sharded_parity(mp, dist, torch, m, rank, local, K_total, T, precision, steps=2):
    """Sharded == single, checked INSIDE the benchmark run: every rank steps its shard `steps` times from X0 (fresh noise
    counters), then rank 0 alone builds ONE engine at K_total with the same seed and steps it from the same state; the two
    nominal sequences (latest_uvec after the shift) and predicted states must agree to summation order."""
    s = X0.copy()
    for _ in range(steps):
        s = m.get_path(s, GOAL)
    U_sh = m.latest_uvec
    rec = None
    if rank == 0:
        one = mp.MPPI(horizon=T, samples=K_total, precision=precision, seed=0, device=local)
        s1 = X0.copy()
        for _ in range(steps):
            s1 = one.get_path(s1, GOAL)
        U1 = one.latest_uvec
        one.close()
        errU = float(np.max(np.abs(U_sh - U1)) / max(np.max(np.abs(U1)), 1e-300))
        errx = float(np.max(np.abs(s - s1)))
        rec = {"against": "one engine at K_total=%d on rank 0, same seed / noise counters, %d closed-loop steps" % (K_total, steps),
               "max_rel_err_U": errU, "max_abs_err_x_next": errx, "tol_rel_U": 1e-9, "pass": bool(errU < 1e-9 and errx < 1e-11)}
    dist.barrier()
    # back to step 0 of the noise stream and a zero nominal on every rank
    m.mppi.use_philox(0)
    m.initialize()
    return rec


def _bench_sharded(args, dist, torch, m, steps):
    """device-timed closed loop of a sharded engine (max over ranks is taken by the caller)"""
    if m.exchange == "p2p":
        # the exchange lives inside the reduce kernel: time the device-resident closed loop with CUDA events on the
        # engine's launch stream (ranks run in lockstep through the rows' flags)
        m.mppi.goal = GOAL
        dist.barrier()
        torch.cuda.synchronize()
        r = m.mppi.bench(X0, steps=steps, warmup=args.warmup, flush_l2=True, per_kernel=False)
        return r["step_ms"], r["launches"], "CUDA events on the engine's launch stream around the two launches of each step, max over ranks", r
    s = X0.copy()
    for _ in range(args.warmup):
        s = m.get_path(s, GOAL)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    st = m.stream if m.stream is not None else torch.cuda.current_stream()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    dist.barrier()
    torch.cuda.synchronize()
    for i in range(steps):
        with torch.cuda.stream(st):
            flush.fill_(i & 0xff)
            ev[i][0].record(st)
        s = m.get_path(s, GOAL)
        ev[i][1].record(st)
    torch.cuda.synchronize()
    return (sum(a.elapsed_time(b) for a, b in ev) / steps, 3 * steps,
            "CUDA events on the stream shared by the engine kernels and the NCCL all-gather, max over ranks", None)


def run_multi(args):
    import torch
    import torch.distributed as dist
    import motion_planning_b200 as mp
    from motion_planning_b200 import _capi
    from motion_planning_b200.distributed import ShardedMPPI
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    # rank 0 prints ONE JSON line on stdout: everything the libraries write there (NCCL's version banner is printed at
    # every NCCL_DEBUG level but NONE) goes to stderr instead; the JSON line is written to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K_total, T = K_PER_GPU * world, T_HORIZON
    m = ShardedMPPI(T, K_total, precision=args.precision, seed=0, device=local, exchange=args.exchange)
    lib, h = m.mppi._lib, m.mppi._h
    parity = _sharded_parity(mp, dist, torch, m, rank, local, K_total, T, args.precision)
    sampler = ClockSampler(local)         # every rank samples ITS GPU (rank 0's goes into `clocks`, all medians into the line)
    dev_ms, launches, timing, _ = _bench_sharded(args, dist, torch, m, args.steps)
    # e2e: host x0 in, (u, x_next) out every step through the C ABI, closed loop on the host; every timed call starts from a
    # flushed L2 on every rank (as at N=1), the flush outside the timed intervals
    m.initialize()
    x, u, xn = X0.copy(), np.empty(2), np.empty(3)
    if m.exchange == "p2p":
        px, pu, pn = _capi.dptr(x), _capi.dptr(u), _capi.dptr(xn)
        _capi.check(lib.mppi_set_goal(h, _capi.dptr(GOAL)), "mppi_set_goal")

        def one():
            if lib.mppi_step(h, px, pu, pn) != 0:
                raise RuntimeError("mppi_step failed: " + lib.mppi_last_error().decode())
            x[:] = xn
    else:
        def one():
            x[:] = m.get_path(x.copy(), GOAL)
    for _ in range(args.warmup):
        one()
    dist.barrier()
    torch.cuda.synchronize()
    acc = 0.0
    for _ in range(args.steps):
        _capi.check(lib.mppi_debug_flush_l2(h), "mppi_debug_flush_l2")
        t0 = time.perf_counter()
        one()
        acc += time.perf_counter() - t0
    torch.cuda.synchronize()
    e2e_ms = acc / args.steps * 1e3
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    launch_info = m.mppi.launch_info()
    io = m.mppi.io_bytes()
    exchange = m.exchange
    m.mppi.close()

    # ---- BASELINE.json configs[4]: K_total = 2097152, T = 128 sharded over the ranks of this run (strong scaling) ----
    c5 = None
    if not args.no_config5:
        m5 = ShardedMPPI(T5, K5_TOTAL, precision=args.precision, seed=0, device=local, exchange=args.exchange)
        parity5 = _sharded_parity(mp, dist, torch, m5, rank, local, K5_TOTAL, T5, args.precision)
        steps5 = max(3, min(args.steps, 30))
        ms5, launches5, _, _ = _bench_sharded(args, dist, torch, m5, steps5)
        t5 = torch.tensor([ms5], dtype=torch.float64, device="cuda")
        dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        ms5 = float(t5[0])
        info5 = m5.mppi.launch_info()
        m5.mppi.close()
        # every rank's shard timed ALONE (no exchange, world_size 1, same K per GPU): separates what the GPUs differ by from what
        # the coupling costs -- the coupled step cannot be faster than the slowest shard
        alone = mp.MPPI(horizon=T5, samples=K5_TOTAL // world, precision=args.precision, seed=0, device=local)
        alone.goal = GOAL
        alone_ms = alone.bench(X0, steps=max(3, min(steps5, 10)), warmup=3, flush_l2=True, per_kernel=False)["step_ms"]
        alone.close()
        ta = torch.zeros(world, dtype=torch.float64, device="cuda")
        ta[rank] = alone_ms
        dist.all_reduce(ta, op=dist.ReduceOp.SUM)
        alone_all = [float(v) for v in ta.cpu()]
        n1_ms = None
        if rank == 0:     # the same workload on ONE GPU, timed in this run: the base of the strong-scaling curve
            one5 = mp.MPPI(horizon=T5, samples=K5_TOTAL, precision=args.precision, seed=0, device=local)
            one5.goal = GOAL
            n1_ms = one5.bench(X0, steps=max(3, min(steps5, 10)), warmup=3, flush_l2=True, per_kernel=False)["step_ms"]
            one5.close()
        dist.barrier()
        c5 = {"workload": "diff-drive parallel-park K_total=%d T=%d (BASELINE.json configs[4]) sharded over %d GPUs, %d rollouts "
                          "per GPU, one %d-byte record exchanged per step" % (K5_TOTAL, T5, world, K5_TOTAL // world, T5 * 48),
              "n_gpus": world, "scaling": "strong", "ms_per_step": ms5, "value": K5_TOTAL / (ms5 * 1e-3), "unit": "rollouts/s",
              "state_steps_per_s": K5_TOTAL * T5 / (ms5 * 1e-3), "steps": steps5, "gpu_launches": launches5,
              "one_gpu_same_run_ms_per_step": n1_ms, "shard_alone_ms_per_rank": alone_all, "launch": info5, "parity": parity5}
    clocks = sampler.stop()
    tc = torch.zeros(world, dtype=torch.float64, device="cuda")
    tc[rank] = float(clocks.get("sm_mhz") or 0.0)
    dist.all_reduce(tc, op=dist.ReduceOp.SUM)
    clocks["sm_mhz_per_rank"] = [float(v) for v in tc.cpu()]
    if rank == 0:
        xbytes = T * 48
        line = {
            "metric": METRIC, "value": K_total / (dev_ms * 1e-3), "unit": "rollouts/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"mixed": "f32 rollouts + f64 re-evaluation of the softmin support", "f32": "f32", "f64": "f64"}[args.precision],
            "data": "synthetic",
            "config": {"workload": "diff-drive parallel-park K=%d per GPU (K_total=%d) T=%d, rollouts sharded over ranks, "
                                   "one %d-byte record exchanged all-to-all per step" % (K_PER_GPU, K_total, T, xbytes),
                       "K_total": K_total, "T": T, "precision": args.precision,
                       "exchange": {"p2p": "fused: every reduce block stores its row of the record into every peer's memory over NVLink (CUDA IPC) "
                                           "as flag-in-data words, merges the peers' rows of its own t, and the finalizer block reads T "
                                           "merged rows; two kernels per rank, no collective call",
                                    "nccl": "ncclAllGather of the device-resident records (torch.distributed)",
                                    "host": "host-staged all-gather"}[exchange],
                       "l2": "flushed between timed steps", "timing": timing, "launch": launch_info},
            "state_steps_per_s": K_total / (dev_ms * 1e-3) * T,
            "e2e": {"value": K_total / (e2e_ms * 1e-3), "unit": "rollouts/s", "h2d_bytes_per_step": io[0] * world,
                    "d2h_bytes_per_step": io[1] * world, "ms_per_step": e2e_ms,
                    "l2": "flushed on every rank before every timed call (outside the timed interval)"},
            "gpu_launches": launches, "clocks": clocks, "parity": parity, "config5": c5,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="mixed", choices=["mixed", "f32", "f64"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-config5", action="store_true", help="skip the K=2097152 T=128 block (BASELINE.json configs[4])")
    ap.add_argument("--ref-full", action="store_true", help="--impl reference: time EVERY step at the full K=65536 (15-40 s per step)")
    ap.add_argument("--no-ref-full-step", dest="ref_full_step", action="store_false",
                    help="--impl reference: skip the single full-K step that checks the sampled figure")
    ap.add_argument("--exchange", default=None, choices=["p2p", "nccl", "host"], help="multi-GPU record exchange (default p2p)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return run_multi(args)
    if args.gpus > 1:
        sys.exit("launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node %d --master-addr 127.0.0.1 "
                 "bench.py --gpus %d ..." % (args.gpus, args.gpus))
    return run_single(args)


if __name__ == "__main__":
    main()
