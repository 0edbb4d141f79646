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


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The Python harness mirrors the C structs by hand: compile the header with gcc and compare sizes
    and the offsets of the last members, so that a drifted field is caught without a GPU."""
    import ctypes as C
    import subprocess
    import cofusion_b200 as cfb
    src = tmp_path / "sizes.c"
    src.write_text('''
#include <stddef.h>
#include <stdio.h>
#include "cofusion_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(cfb_track_stats), sizeof(cfb_track_params),
         sizeof(cfb_seg_params), sizeof(cfb_model_data), sizeof(cfb_cofusion_params),
         offsetof(cfb_cofusion_params, seg), offsetof(cfb_cofusion_params, modelSpawnOffset),
         offsetof(cfb_model_data, left), offsetof(cfb_track_stats, so3_iterations));
  return 0;
}
''')
    exe = tmp_path / "sizes"
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(cfb.TrackStats), C.sizeof(cfb.TrackParams), C.sizeof(cfb.SegParams), C.sizeof(cfb.ModelData),
            C.sizeof(cfb.CoFusionParams), cfb.CoFusionParams.seg.offset, cfb.CoFusionParams.modelSpawnOffset.offset,
            cfb.ModelData.left.offset, cfb.TrackStats.so3_iterations.offset]
    assert got == want, (got, want)


def _build_binding(tmp_path):
    import subprocess
    exe = tmp_path / "binding_test"
    libdir = os.path.join(ROOT, "cofusion_b200")
    subprocess.run(["g++", "-std=c++14", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "binding", "cofusion_binding.cpp"), "-L", libdir, "-lcofusion_b200",
                    "-Wl,-rpath," + libdir, "-o", str(exe)], check=True)
    return exe


def test_reference_side_binding_compiles_and_fails_loudly_without_a_device(tmp_path):
    """INTEGRATION.md section 1 as a compiled C++ translation unit (tests/binding/cofusion_binding.cpp): the
    header is valid C++ and C, the binding links against the shared library alone, and without a CUDA device the
    reference-side constructor throws the library's error instead of computing on the CPU."""
    import subprocess
    cfb_build.build()
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c",
                    "-include", "cofusion_b200.h", "/dev/null"], check=True)
    exe = _build_binding(tmp_path)
    if cofusion_b200.lib().cfb_device_count() > 0:
        return  # the GPU suite runs it for real (tests/test_boundary_gpu.py)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "no CUDA device" in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_frame_struct_matches_the_header(tmp_path):
    import ctypes as C
    import subprocess
    import cofusion_b200 as cfb
    src = tmp_path / "frame.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "cofusion_b200.h"\nint main(void){printf("%zu %zu %zu\\n", '
                   'sizeof(cfb_frame), offsetof(cfb_frame, timestamp), offsetof(cfb_frame, mask));return 0;}\n')
    exe = tmp_path / "frame"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got == [C.sizeof(cfb.Frame), cfb.Frame.timestamp.offset, cfb.Frame.mask.offset], got
