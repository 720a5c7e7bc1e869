"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference MPPI module.

The reference hot path is the extension-less Python script
``/root/reference/control/src/mppi`` (class ``MPPI`` at control/src/mppi:61-213).
It imports rospy / tf / matplotlib / nav_msgs / geometry_msgs at module level
(control/src/mppi:4-13), none of which exist in this image, so empty stub
modules are planted in ``sys.modules`` before the file is executed.  Nothing of
the hot path touches those modules.

Two sources, tried in this order:
  1. ``/root/reference/control/src/mppi``  (only exists in the build container)
  2. ``oracle/_ref/mppi.pyc``  -- byte-code compiled FROM the sources where they
     lie by ``oracle/build_ref.py`` (git-ignored build product, travels to the
     GPU box with the snapshot; no reference source text is copied).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this file.  The product (motion_planning_b200/) never does.
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/control/src/mppi"
REF_PYC = os.path.join(_HERE, "_ref", "mppi.pyc")

_STUBS = {
    "rospy": {},
    "tf": {},
    "matplotlib": {},
    "matplotlib.pyplot": {},
    "matplotlib.animation": {},
    "nav_msgs": {},
    "nav_msgs.msg": {"Odometry": type("Odometry", (), {})},
    "geometry_msgs": {},
    "geometry_msgs.msg": {
        "Twist": type("Twist", (), {}),
        "Quaternion": type("Quaternion", (), {}),
        "Vector3": type("Vector3", (), {}),
    },
}


def _plant_stubs():
    planted = []
    for name, attrs in _STUBS.items():
        if name in sys.modules:
            continue
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        planted.append(name)
    # "from matplotlib import animation" needs the attribute on the parent
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    return planted


def available():
    """Which flavour of the real reference can be loaded here (or None)."""
    if os.path.exists(REF_SRC):
        return "source"
    if os.path.exists(REF_PYC):
        return "pyc"
    return None


def load_reference(fresh=True):
    """Return the reference module object (attributes MPPI, rk4, dd_dynamics...).

    NB: executing the module runs ``np.random.seed(0)`` (control/src/mppi:15).
    """
    kind = available()
    if kind is None:
        raise RuntimeError("reference MPPI not available (no %s, no %s)" % (REF_SRC, REF_PYC))
    planted = _plant_stubs()
    try:
        name = "ref_mppi"
        if fresh and name in sys.modules:
            del sys.modules[name]
        if kind == "source":
            loader = importlib.machinery.SourceFileLoader(name, REF_SRC)
        else:
            loader = importlib.machinery.SourcelessFileLoader(name, REF_PYC)
        spec = importlib.util.spec_from_loader(name, loader)
        mod = importlib.util.module_from_spec(spec)
        old = sys.dont_write_bytecode
        sys.dont_write_bytecode = True          # never write next to the read-only reference
        try:
            loader.exec_module(mod)
        finally:
            sys.dont_write_bytecode = old
        mod.__ref_kind__ = kind
        return mod
    finally:
        for n in planted:                      # do not leak fake rospy etc. into the process
            sys.modules.pop(n, None)
