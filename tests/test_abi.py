"""CPU checks of the boundary: the C-ABI library loads and exports every symbol include/*.h declares
(no compute calls without a GPU), and the ctypes signature table covers exactly that set."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cppf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cppf_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("cppf_vote_center", "cppf_grid_argmax", "cppf_backvote_filter", "cppf_rotation_hist",
                 "cppf_shot_compute", "cppf_estimate_normal", "cppf_heads_forward", "cppf_decode_targets"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from cppf2_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/cppf_b200.h but not exported: {missing}"


def test_signature_table_matches_header():
    from cppf2_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.cppf_version() >= 100
    assert lib.cppf_error_string(1).decode() == "invalid argument"
    assert lib.cppf_sphere_band(720, 0.99939084) == 15


def test_struct_layouts_match_header_sizes(tmp_path):
    """ctypes mirrors of the ABI structs have the size the C compiler gives the header's structs."""
    import subprocess
    from cppf2_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include "cppf_b200.h"\n#include <stdio.h>\nint main(void){printf("%zu %zu %zu %zu\\n", '
                   'sizeof(cppf_grid_geom), sizeof(cppf_center), sizeof(cppf_backvote_summary), sizeof(cppf_pose));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_lib.GridGeom), ctypes.sizeof(_lib.Center), ctypes.sizeof(_lib.BackvoteSummary),
                     ctypes.sizeof(_lib.Pose)]


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from cppf2_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.CppfError):
        _lib.load()
