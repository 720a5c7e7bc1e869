"""TEST INFRASTRUCTURE ONLY -- "compile" the Python reference into oracle/_ref/.

The reference hot path is an interpreted Python script, so the analogue of
"compile the reference's own few source files into oracle/_ref/*.so" is
byte-compiling /root/reference/control/src/mppi (where it lies, unmodified)
into oracle/_ref/mppi.pyc.  The .pyc is a build product: git-ignored, but it
travels to the GPU box with the gpurun snapshot so that the *real* reference
can be used there as the checker and timed as the CPU baseline
(bench.py --impl reference).  No reference source text enters the repo.
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/control/src/mppi"
OUT = os.path.join(HERE, "_ref", "mppi.pyc")


def build(quiet=False):
    if not os.path.exists(SRC):
        if not quiet:
            print("oracle/build_ref: %s absent (GPU box?) -- keeping prebuilt %s" % (SRC, OUT))
        return os.path.exists(OUT)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    py_compile.compile(SRC, cfile=OUT, dfile="reference:control/src/mppi", doraise=True,
                       invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    if not quiet:
        print("oracle/build_ref: wrote", OUT)
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
