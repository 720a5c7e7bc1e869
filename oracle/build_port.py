"""TEST INFRASTRUCTURE ONLY -- build oracle/mppi_port.c into oracle/_build/libmppi_port.so (gcc, OpenMP)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "mppi_port.c")
OUT = os.path.join(HERE, "_build", "libmppi_port.so")


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # -ffp-contract=off: keep the reference's operation order (no fused multiply-adds)
    cmd = ["gcc", "-O2", "-march=x86-64-v2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", SRC, "-o", OUT, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("gcc failed")
    print("oracle/build_port: wrote", OUT)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
