"""CPU: the C-ABI library builds, loads and exports every symbol include/mppi_b200.h declares;
the product fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re
import subprocess

import pytest

from motion_planning_b200 import _capi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "mppi_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"MPPI_API\s+[\w\s\*]+?\b(mppi_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == _capi.exported_symbols()


def test_library_exports_every_declared_symbol():
    lib = _capi.load()                      # raises ImportError if the extension is not built
    for name in declared_symbols():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (mppi_\w+)", out))
    assert exported == set(declared_symbols())


def test_params_struct_matches_header_size():
    lib = _capi.load()
    p = _capi.MppiParams()
    assert lib.mppi_default_params(ctypes.byref(p)) == 0
    assert p.struct_size == ctypes.sizeof(_capi.MppiParams)
    assert p.abi_version == _capi.ABI_VERSION
    # reference constants, control/src/mppi:18-20,62-73,88-89
    assert (p.K, p.T) == (10, 100)
    assert list(p.q) == [1e3, 1e3, 0.0] and list(p.p1) == [1e3, 1e3, 1e3]
    assert p.lambda_ == 1e-3 and p.sig[0] == 0.9 and p.u_max[0] == 6.35492
    assert p.wheel_radius == 0.033 and p.wheel_base == 0.16 and p.eps_floor == 1e-8


def test_sass_is_sm100a_with_tma_bulk_copy():
    """The rollout kernel stages the nominal block with a TMA bulk copy (SASS: UBLKCP) and reduces the
    floor sums with REDUX; nothing may be compiled for another arch."""
    r = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout
    assert not re.search(r"sm_(?!100a)\d+", r.stdout)


def test_no_cpu_fallback_without_gpu():
    lib = _capi.load()
    if lib.mppi_device_count() > 0:
        pytest.skip("a GPU is present")
    import motion_planning_b200 as mp
    with pytest.raises(mp.MppiError) as ei:
        mp.MPPI(horizon=32, samples=128)
    assert ei.value.status == 3      # MPPI_ERR_NO_DEVICE


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under motion_planning_b200/ may import or load it."""
    pkg = os.path.join(ROOT, "motion_planning_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert "mppi_oracle" not in txt and "ref_loader" not in txt, f


def test_every_struct_layout_matches_its_ctypes_mirror(tmp_path):
    """include/mppi_b200.h compiled as plain C (the header is the contract a cgo / JNI / ctypes binding reads): sizeof and every
    field offset of mppi_params, mppi_user_model and mppi_timing against the ctypes mirrors the Python host side uses."""
    structs = {"mppi_params": _capi.MppiParams, "mppi_user_model": _capi.MppiUserModel, "mppi_timing": _capi.MppiTiming}
    lines = ["#include <stddef.h>", "#include <stdio.h>", '#include "mppi_b200.h"', "int main(void) {"]
    for cname, mirror in structs.items():
        lines.append('  printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in mirror._fields_:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname.rstrip("_")))
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = str(tmp_path / "layout")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, fname, value = line.split()
        mirror = structs[cname]
        want = ctypes.sizeof(mirror) if fname == "sizeof" else getattr(mirror, fname).offset
        assert int(value) == want, line
        seen += 1
    assert seen == sum(len(m._fields_) + 1 for m in structs.values())
