"""The C-ABI library loads and exports every symbol include/manette_b200.h declares (no compute calls: those
need a GPU and live in the -m gpu tests)."""
import ctypes as C
import os
import re

import util


def _declared_symbols():
    text = open(os.path.join(util.ROOT, "include", "manette_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from manette_b200 import build, _native
    path = build.build()
    lib = C.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), "libmanette_b200.so does not export %s" % name
    # and the Python binding covers exactly the header
    assert sorted(_native.SYMBOLS) == declared


def test_library_is_built_for_sm_100a_only():
    import subprocess
    from manette_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_a_gpu():
    """The product must fail loudly, not fall back to the oracle, when there is no CUDA device."""
    import pytest
    import torch
    import manette_b200 as mb
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mb._native.NativeError):
        mb.DevicePool([("pong", util.rom_bytes("pong"), 2)])


def test_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under manette_b200/ may import, include or load it."""
    pkg = os.path.join(util.ROOT, "manette_b200")
    bad = re.compile(r"(import\s+(host_path|orc_loader|ref_harness)|from\s+oracle|#include\s+\".*oracle|liborc|sys\.path.*oracle)")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not bad.search(src), f


def test_start_noop_schedule_is_a_pure_function():
    from manette_b200 import start_noops
    vals = [start_noops(3, e, ep) for e in range(50) for ep in range(20)]
    assert min(vals) >= 0 and max(vals) <= 30 and len(set(vals)) > 20
    assert vals == [start_noops(3, e, ep) for e in range(50) for ep in range(20)]
