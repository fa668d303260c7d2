"""CPU checks of the drop-in boundary: the shared library loads without a GPU, exports every symbol include/dsc.h
declares, refuses to work without a device (no CPU fallback), and the product never imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "dsc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dsc_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(pkg):
    lib = pkg.load_library()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dsc.h but not exported by libdsc_b200.so"
    assert sorted(pkg.binding.EXPORTS) == names


def test_no_cpu_fallback(pkg):
    import subprocess
    import sys
    code = ("import __graft_entry__ as g, ctypes; p=g.package(); lib=p.load_library(); h=ctypes.c_void_p();"
            "print(lib.dsc_create(0, ctypes.byref(h)))")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, stdout=subprocess.PIPE, text=True).stdout.strip()
    assert out == "-3"        # DSC_ERR_NO_DEVICE


def test_status_strings_and_version(pkg):
    lib = pkg.load_library()
    assert lib.dsc_version() >= 100
    assert b"no CPU fallback" in lib.dsc_status_string(-3)
    assert lib.dsc_status_string(0) == b"ok"


def test_struct_sizes_match_header(pkg):
    b = pkg.binding
    assert ctypes.sizeof(b.Camera) == 36 and ctypes.sizeof(b.Pair) == 168
    assert ctypes.sizeof(b.TriParams) == 24 and ctypes.sizeof(b.Weights) == 48
    assert ctypes.sizeof(b.PcgParams) == 16 and ctypes.sizeof(b.IterRecord) == 40 and ctypes.sizeof(b.OptStats) == 72


def test_product_does_not_touch_the_oracle():
    pkg_dir = os.path.join(ROOT, "triangulation-in-deformable-scenes_b200")
    for dp, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".cpp", ".h", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "oracle/" not in txt.replace("oracle/f32.py", "").replace("oracle/graph.py", ""), f
