"""TEST INFRASTRUCTURE ONLY -- compile the reference's OWN occupancy-grid sources where they lie
(/root/reference/map/src/map/{map,grid,prm}.cpp, headers under /root/reference/map/include) into oracle/_ref/libmapref.so.

The map library's third-party dependencies (rigid2d, nuslam from nuturtle.rosinstall:1-6; Eigen 3) are not in this image; the
few symbols the three files use from them are provided by minimal stand-in headers under oracle/ref_stubs/ (ours, documented
there).  The reference's build system (catkin / cmake) is NOT run.  No reference source is copied into the repo: the compiler
reads the files from /root/reference, the output goes to the git-ignored oracle/_ref/ (which travels to the GPU box).

    python -m oracle.build_map_ref
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_MAP = "/root/reference/map"
OUT = os.path.join(HERE, "_ref", "libmapref.so")


def build(quiet=True):
    srcs = [os.path.join(REF_MAP, "src", "map", f) for f in ("map.cpp", "grid.cpp", "prm.cpp")]
    if not all(os.path.exists(s) for s in srcs):
        return OUT if os.path.exists(OUT) else None      # GPU box: only the prebuilt file exists
    gxx = shutil.which("g++")
    if not gxx:
        return None
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    shim = os.path.join(HERE, "map_ref_shim.cpp")
    deps = srcs + [shim] + [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(HERE, "ref_stubs")) for f in fs]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    # -O1 -ffp-contract=off: plain IEEE double arithmetic, no FMA contraction (the catkin build is -O2 on x86-64 without
    # -march flags: same arithmetic); -include cmath/iostream: the reference relies on transitive includes of its dependencies
    cmd = [gxx, "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-w",
           "-I", os.path.join(HERE, "ref_stubs"), "-I", os.path.join(REF_MAP, "include"),
           "-include", "cmath", "-include", "iostream", "-include", "stdexcept", "-include", "algorithm", "-include", "random",
           "-o", OUT] + srcs + [shim]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("oracle/build_map_ref: g++ failed")
    if not quiet:
        print("oracle/build_map_ref: wrote", OUT)
    return OUT


if __name__ == "__main__":
    print(build(quiet=False))
