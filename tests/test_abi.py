"""CPU-side checks of the C-ABI library: it loads and exports every symbol include/apss.h declares.
No compute calls here (no GPU in this container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "apss.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(apss_[a-z_]+)\s*\(", txt)))


def test_header_symbols_match_binding_list():
    import apss_b200
    assert _declared_symbols() == sorted(apss_b200.native.EXPORTS)


def test_library_loads_and_exports_every_declared_symbol():
    import apss_b200
    path = apss_b200.native.LIB_PATH
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(path)
    for sym in _declared_symbols():
        assert hasattr(lib, sym), sym
    lib.apss_abi_version.restype = ctypes.c_int32
    assert lib.apss_abi_version() == apss_b200.native.ABI_VERSION


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C layout (checked against a tiny C program compiled on the fly)."""
    import subprocess
    import tempfile
    import apss_b200
    src = '#include "apss.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(apss_config), sizeof(apss_batch_result), sizeof(apss_stats));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        out = subprocess.check_output([os.path.join(d, "s")]).decode().split()
    n = apss_b200.native
    assert [int(x) for x in out] == [ctypes.sizeof(n.Config), ctypes.sizeof(n.BatchResultC), ctypes.sizeof(n.StatsC)]


def test_no_cpu_fallback_without_device():
    """On a box without CUDA the product path must fail loudly, not fall back."""
    import torch
    import apss_b200
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(apss_b200.native.ApssError) as ei:
        apss_b200.native.Index(1024, 0.5)
    assert ei.value.code == -6


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "all-pairs-similarity_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "apss_oracle" not in txt, f
