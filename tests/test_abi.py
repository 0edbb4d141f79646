"""-m "not gpu": the C-ABI library builds, loads and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import cofusion_b200
from cofusion_b200 import build as cfb_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "cofusion_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cfb_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    path = cfb_build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, "declared in include/cofusion_b200.h but not exported: %s" % missing


def test_no_cpu_fallback_without_device():
    """Without a GPU the module must fail loudly, not compute on the CPU."""
    lib = cofusion_b200.lib()
    if lib.cfb_device_count() > 0:
        return
    h = ctypes.c_void_p()
    rc = lib.cfb_odom_create(640, 480, ctypes.c_float(320), ctypes.c_float(240), ctypes.c_float(528),
                             ctypes.c_float(528), ctypes.c_float(0.1), ctypes.c_float(0.34), ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.cfb_last_error()


def test_product_never_touches_the_oracle():
    """The product sources must not include, link, import or dlopen anything under oracle/
    (comments may cite it)."""
    pkg = os.path.join(ROOT, "cofusion_b200")
    bad = re.compile(r"(#\s*include[^\n]*oracle|liboracle|libcfref|cf_oracle\.h|import\s+orc\b|from\s+orc\b|"
                     r"CDLL\([^)]*oracle|dlopen\([^)]*oracle)")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".cu", ".cuh", ".h", ".py", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert not bad.search(src), f
