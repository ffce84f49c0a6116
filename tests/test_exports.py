"""The C-ABI libraries load without a GPU and export every symbol include/*.h declares
(no compute calls here: those are the -m gpu tests)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECL = re.compile(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(ibis_b200_\w+)\s*\(", re.M)


def declared(header):
    with open(os.path.join(ROOT, "include", header)) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)       # comments mention symbols too
    names = set(DECL.findall(text))
    return sorted(n for n in names if not n.endswith("_config") and not n.endswith("_desc"))


@pytest.mark.parametrize("header,lib", [("ibis_b200.h", "libibis_b200.so"),
                                        ("ibis_b200_host.h", "libibis_b200_host.so")])
def test_every_declared_symbol_is_exported(header, lib):
    path = os.path.join(ROOT, "ibis_b200", "lib", lib)
    assert os.path.exists(path), f"{lib} is not built (python __graft_entry__.py)"
    L = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    names = declared(header)
    assert len(names) >= 8
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"{lib} does not export {missing}"


def test_ctypes_mirror_covers_the_header():
    """ibis_b200/lib.py declares argtypes for every entry point of ibis_b200.h."""
    with open(os.path.join(ROOT, "ibis_b200", "lib.py")) as f:
        src = f.read()
    missing = [n for n in declared("ibis_b200.h") if n not in src]
    assert not missing, f"ibis_b200/lib.py does not bind {missing}"


def test_product_has_no_cpu_fallback():
    """The package never imports the oracle or the host emulation build."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ibis_b200")):
        for fn in files:
            if not fn.endswith((".py", ".cu", ".cpp", ".h", ".hpp")):
                continue
            with open(os.path.join(dirpath, fn)) as f:
                text = f.read()
            assert "import refharness" not in text and "from oracle" not in text and "libibis_ref" not in text, fn
            assert "import restatement" not in text, fn
            assert "ibis_b200_emu" not in text or fn in ("abi.cu", "backend.h"), fn


@pytest.mark.parametrize("header", ["ibis_b200.h", "ibis_b200_host.h"])
def test_headers_are_plain_c(header):
    """the boundary is a C ABI: the headers compile as C99 (no C++ types, no torch types in any signature)"""
    import subprocess
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-x", "c", "-fsyntax-only", "-I", ROOT, "-"],
                       input=f'#include "include/{header}"\n', text=True, capture_output=True)
    assert r.returncode == 0, r.stderr

