"""The C-ABI shared library loads without a GPU and exports every symbol include/pcab200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "pcab200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcab_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pcaccumulation_b200 import _lib

    lib = _lib.lib()
    names = header_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == names


def test_version_and_error_string():
    from pcaccumulation_b200 import _lib

    lib = _lib.lib()
    assert lib.pcab_version() >= 100
    assert isinstance(lib.pcab_last_error(), bytes)


def test_workspace_queries_are_pure_host_calls():
    from pcaccumulation_b200 import _lib

    assert _lib.size("pcab_chamfer_workspace", _lib.I(2), _lib.I(100), _lib.I(300)) >= 2 * 300 * 8
    assert _lib.size("pcab_ego_pairs_workspace", _lib.I(4)) > 4 * 1024 * 1024 * 4
    assert _lib.lib().pcab_pfn_pack_size() == 9 * 64 + 64 + 3 * (64 * 32 + 32 + 32 * 32 + 32 + 64 * 32) + 32 * 32 + 32


def test_product_path_has_no_oracle_or_cpu_fallback():
    pkg = os.path.join(ROOT, "pcaccumulation_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_header_is_plain_c_and_workspaces_return_size_t(tmp_path):
    """include/pcab200.h is what a C host (or a cgo / JNI stub) includes: it must compile as C99 without CUDA or torch headers,
    and every *_workspace query must be bound with a size_t return type (a default int return truncates above 2 GB)."""
    import shutil
    import subprocess

    from pcaccumulation_b200 import _lib

    cc = shutil.which("gcc") or shutil.which("cc")
    assert cc, "a C compiler is part of the build environment"
    src = tmp_path / "abi_check.c"
    src.write_text('#include "pcab200.h"\nint main(void) { return pcab_version() > 0 ? 0 : 1; }\n')
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = _lib.lib()
    for name in _lib.SYMBOLS:
        if name.endswith("_workspace"):
            assert getattr(lib, name).restype is ctypes.c_size_t, name
    # two of the newer queries, as pure host calls
    assert _lib.size("pcab_nn_workspace", _lib.I(1000), _lib.I(350_000)) > 350_000 * 16
    assert _lib.size("pcab_seg_loss_workspace", _lib.L(415_000)) > 415_000 * 4 * 8
