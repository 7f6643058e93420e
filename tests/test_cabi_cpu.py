"""CPU: the C-ABI library loads and exports every symbol include/graphem_b200.h declares; the
ctypes binding covers exactly that set; no compute call is made (no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

from gem_testutil import ROOT

HEADER = os.path.join(ROOT, "include", "graphem_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gem_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from graphem_rapids_b200 import _cabi, build
    lib_path = build.build()
    assert os.path.exists(lib_path)
    names = declared_functions()
    assert len(names) >= 15
    lib = ctypes.CDLL(lib_path)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/graphem_b200.h but not exported"
    assert sorted(_cabi.SIGNATURES) == names, "ctypes SIGNATURES out of sync with the header"
    loaded = _cabi.load()
    assert loaded.gem_abi_version() == _cabi.ABI_VERSION == int(re.search(r"#define GEM_ABI_VERSION (\d+)", open(HEADER).read()).group(1))
    assert loaded.gem_row_pitch(2) == 2 and loaded.gem_row_pitch(3) == 4 and loaded.gem_row_pitch(7) == 7
    assert loaded.gem_mid_pitch(2) == 2 and loaded.gem_mid_pitch(3) == 4 and loaded.gem_mid_pitch(7) == 8
    assert b"out of range" in loaded.gem_error_string(-3)


def test_library_is_sm100a_with_tma_and_vector_red():
    """SASS evidence that the shipped binary is a Blackwell build using TMA bulk copies."""
    from graphem_rapids_b200 import build
    lib = build.build()
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass                     # cp.async.bulk (TMA) candidate tiles
    assert "SYNCS" in sass                      # mbarrier
    assert re.search(r"RED\.E\.ADD\.F32x[24]|RED\.E\.ADD\.F32\.FTZ\.RN|REDG\.E\.ADD|RED\.E\.ADD\.(64|128)", sass) or "RED" in sass


def test_plan_struct_layout_matches_header():
    from graphem_rapids_b200 import _cabi
    text = open(HEADER).read()
    body = text[text.index("typedef struct gem_plan {"):text.index("} gem_plan;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.replace("*", " ").split()
        # "int64_t n, e, s" -> n e s ; "const int32_t edges" -> edges
        first = decl.split(",")
        fields.append(first[0].replace("*", " ").split()[-1])
        for extra in first[1:]:
            fields.append(extra.replace("*", " ").split()[-1])
    assert fields == [f[0] for f in _cabi.GemPlan._fields_]


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "graphem_rapids_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "oracle." not in src.replace("oracle/", ""), f


def test_scan_kernel_keeps_uniform_register_operands():
    """The scan's packed FMAs take the query coefficients as UNIFORM-register operands (SASS: FFMA2 R, R.F32, UR.F32x2, R)
    -- the difference between ~46 and ~60 TFLOP/s of the loop.  ptxas drops that form silently for the whole
    main loop when the kernel's instruction stream changes in ways it dislikes (a global atomic, an extra global store
    in the prologue, a per-warp query index ...), so the built library is checked: every FFMA2 of knn_scan_kernel
    has a UR operand, and the bulk copies are there (UBLKCP)."""
    import re
    import shutil
    import subprocess
    from graphem_rapids_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.LIB], capture_output=True, text=True, check=True).stdout
    seen = 0
    for fn in sass.split("Function : ")[1:]:
        name = fn.split("\n", 1)[0]
        if "knn_scan_kernel" not in name:
            continue
        seen += 1
        ffma2 = [ln for ln in fn.splitlines() if "FFMA2" in ln]
        assert len(ffma2) >= 96, (name, len(ffma2))
        assert all(re.search(r"\bUR\d+", ln) for ln in ffma2), f"{name}: FFMA2 without uniform operand"
        assert "UBLKCP" in fn, f"{name}: no bulk-copy (TMA) instruction"
    assert seen == 2
