"""Build recipe of libmppi_b200.so (hand-written CUDA for sm_100a + the C ABI), in-tree.

    python -m motion_planning_b200.build        # or __graft_entry__.build()

nvcc cross-compiles without a GPU; the five translation units compile in parallel.  The library
links cudart statically and has no other dependency (no torch types cross the ABI).
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "lib", "libmppi_b200.so")
UNITS = ["rollout_f32_softmin.cu", "rollout_f32_screen.cu", "rollout_f64_softmin.cu", "reduce.cu", "engine.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), tag=None):
    """tag/defines: build an experimental variant into lib/libmppi_b200_<tag>.so (select it at run time with
    the environment variable MPPI_B200_LIB); the default build is the product."""
    global OBJ, LIB
    nvcc = _nvcc()
    obj_dir, lib_path = OBJ, LIB
    if tag:
        obj_dir = os.path.join(HERE, "build_" + tag)
        lib_path = os.path.join(HERE, "lib", "libmppi_b200_%s.so" % tag)
    return _build(nvcc, obj_dir, lib_path, force, verbose, list(defines))


def _build(nvcc, OBJ, LIB, force, verbose, defines):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))]
    hdrs.append(os.path.join(HERE, "..", "include", "mppi_b200.h"))
    jobs = []
    for u in UNITS:
        src = os.path.join(CSRC, u)
        obj = os.path.join(OBJ, u[:-3] + ".o")
        if force or _stale(obj, [src] + hdrs):
            jobs.append([nvcc] + FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(5, os.cpu_count() or 1)) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
    objs = [os.path.join(OBJ, u[:-3] + ".o") for u in UNITS]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    tags = [a[6:] for a in sys.argv[1:] if a.startswith("--tag=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, tag=tags[0] if tags else None))
