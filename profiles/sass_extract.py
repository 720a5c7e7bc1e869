"""SASS of the step's dominant kernels out of the shipped library (cuobjdump; runs without a GPU):
    python profiles/sass_extract.py [tag]     ->  profiles/<tag>_sass_<kernel>.txt + an instruction histogram / evidence summary
The summary lists the mnemonics that prove the claimed mechanisms (profiling guide): UBLKCP (TMA bulk copy), SYNCS (mbarrier),
ACQBULK / PREEXIT (programmatic dependent launch), REDUX (warp integer add), FFMA.SAT (clip as a saturate), MUFU (Box-Muller on the
SFU pipe), IMAD.WIDE (Philox), and -- for an elementwise ODE kernel -- the ABSENCE of tensor-core instructions (UTC*MMA, HMMA)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
LIB = os.path.join(ROOT, "motion_planning_b200", "lib", "libmppi_b200.so")
KERNELS = {
    "rollout_lean_sm": "_ZN4mppi22rollout_lean_sm_kernelILi0ELi1ELb0EEEvNS_11RolloutArgsE",     # diff-drive, SCREEN, no grid: BASELINE config 2
    "reduce_screen": "_ZN4mppi20reduce_screen_kernelILi0ELb0EEEvNS_10ReduceArgsE",
}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    summary = []
    for short, sym in KERNELS.items():
        out = subprocess.run(["cuobjdump", "-sass", "-fun", sym, LIB], capture_output=True, text=True).stdout
        path = os.path.join(ROOT, "profiles", "%s_sass_%s.txt" % (tag, short))
        # keep the instruction text, drop the hex encodings (second half of every line / every other line)
        slim = []
        for line in out.splitlines():
            if re.match(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", line):
                continue
            slim.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line))
        if short.startswith("rollout"):      # the listing of the dominant kernel is committed; the others appear in the summary only
            open(path, "w").write("\n".join(l for l in slim if l.strip() and not l.startswith(("Fatbin", "====", "arch =", "code version", "host =", "compile_size", "identifier", "\tcode for"))) + "\n")
        ops = collections.Counter()
        for line in out.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                ops[m.group(1)] += 1
        total = sum(ops.values())
        fam = collections.Counter()
        for op, n in ops.items():
            fam[op.split(".")[0]] += n
        summary.append("== %s  (%s)\n   %d instructions; arch %s" % (short, sym, total, re.search(r"arch = (\S+)", out).group(1)))
        summary.append("   by family: " + ", ".join("%s %d" % kv for kv in fam.most_common(24)))
        ev = {k: sum(n for op, n in ops.items() if op.startswith(k)) for k in
              ("UBLKCP", "SYNCS", "ACQBULK", "PREEXIT", "REDUX", "FFMA.SAT", "MUFU", "IMAD.WIDE", "LDS", "STS", "BAR", "UTC", "HMMA", "LDTM", "DFMA", "DMUL")}
        summary.append("   evidence: " + ", ".join("%s %d" % kv for kv in ev.items()))
    text = "\n".join(summary) + "\n"
    open(os.path.join(ROOT, "profiles", "%s_sass_summary.txt" % tag), "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
