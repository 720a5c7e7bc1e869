"""Markdown tables for profiles/README.md from the committed JSON lines of a round:  python profiles/make_tables.py r02"""
import json
import os
import sys

P = os.path.dirname(os.path.abspath(__file__))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"


def line(path):
    return json.loads(open(os.path.join(P, path)).read().strip().splitlines()[-1])


b = line("%s_bench_n1.json" % R)
ref = line("%s_bench_reference_n1.json" % R)
print("## headline\n")
print("| quantity | value |\n|---|---|")
print("| `value` (flushed L2) | %.3g rollouts/s, %.1f us per step |" % (b["value"], b["ms_per_step"] * 1e3))
print("| warm L2 | %.3g rollouts/s, %.1f us per step |" % (b["warm_l2"]["value"], b["warm_l2"]["ms_per_step"] * 1e3))
print("| `e2e` | %.3g rollouts/s, %.1f us (warm %.1f us; python shim %.1f us) |" % (
    b["e2e"]["value"], b["e2e"]["ms_per_step"] * 1e3, b["e2e"]["warm_l2_ms_per_step"] * 1e3, b["e2e"]["python_shim_ms_per_step"] * 1e3))
print("| kernels (eager, events) | %s |" % {k: round(v * 1e3, 2) for k, v in b["kernels_ms"].items()})
print("| roofline | achieved %.1f TFLOP/s of %.1f = %.3f; traffic %s |" % (b["roofline"]["achieved"], b["roofline"]["peak"], b["roofline"]["frac"], b["roofline"]["traffic"]))
print("| other precisions | %s |" % {p: (round(v["ms_per_step"] * 1e3, 1), round(v["e2e_ms"] * 1e3, 1)) for p, v in b["other_precisions"].items()})
print("| cpu_baseline | %.3g rollouts/s on %d cores (%s) |" % (b["cpu_baseline"]["value"], b["cpu_baseline"]["cores"], b["cpu_baseline"]["kind"]))
print("| reference arm | %.3g rollouts/s (sampled K=%s); full-K step: %s |" % (ref["value"], ref["config"].get("sample_K"), ref["config"].get("full_K_step")))
c5 = b.get("config5") or {}
print("| config5 on 1 GPU | %.1f us per step, %.3g rollouts/s, rollout kernel %.1f us, frac %.3f |" % (
    c5["ms_per_step"] * 1e3, c5["value"], c5["rollout_ms"] * 1e3, c5["roofline_frac_rollout_kernel"]))
print("| refine | %s |" % b["refine"])
print("| clocks | %s |" % b["clocks"])
print("\n## scaling\n")
print("| N | weak us/step | rollouts/s | eff | e2e us | parity relerr | config5 us/step | rollouts/s | strong eff (vs N=1 in the same run) | shard alone us (max) | parity |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
base = b["value"]
for n in (1, 2, 4, 8):
    f = "%s_scale_n%d.json" % (R, n)
    if not os.path.exists(os.path.join(P, f)):
        continue
    d = line(f)
    c = d.get("config5") or {}
    par = d.get("parity") or {}
    p5 = c.get("parity") or {}
    one = c.get("one_gpu_same_run_ms_per_step")
    print("| %d | %.1f | %.3g | %.2f | %.1f | %s | %.1f | %.3g | %s | %s | %s |" % (
        n, d["ms_per_step"] * 1e3, d["value"], d["value"] / (n * base), d["e2e"]["ms_per_step"] * 1e3, par.get("max_rel_err_U"),
        c["ms_per_step"] * 1e3, c["value"], ("%.2f" % (one / (n * c["ms_per_step"]))) if one else "1.00",
        ("%.1f" % (max(c["shard_alone_ms_per_rank"]) * 1e3)) if c.get("shard_alone_ms_per_rank") else "-", p5.get("max_rel_err_U")))
print("\n## sweep\n")
print("| config | precision | us/step | rollout us | rollouts/s | state-steps/s | candidates | max dev | head-room | kernel |")
print("|---|---|---|---|---|---|---|---|---|---|")
for l in open(os.path.join(P, "%s_sweep_configs.jsonl" % R)):
    l = l.strip()
    if not l.startswith("{"):
        continue
    d = json.loads(l)
    la = d["launch"]
    print("| %s | %s | %.1f | %.1f | %.3g | %.3g | %d | %.2g | %.2g | %s/%d x%d |" % (
        d["config"], d["precision"], d["ms_per_step"] * 1e3, d["rollout_ms"] * 1e3, d["rollouts_per_s"], d["state_steps_per_s"],
        d["refine_candidates"], d["refine_max_dev"], d.get("refine_head_room", 0), la["variant"], la["block"], la["ctas_per_sm"]))
