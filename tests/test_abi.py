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
                   'sizeof(cppf_grid_geom), sizeof(cppf_center), sizeof(cppf_backvote_summary), sizeof(cppf_pose));'
                   'printf("%zu %zu\\n", sizeof(cppf_vote_params), sizeof(cppf_vote_buffers));'
                   'printf("%zu %zu %zu\\n", sizeof(cppf_instance_io), sizeof(cppf_frame), sizeof(cppf_scale_select));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_lib.GridGeom), ctypes.sizeof(_lib.Center), ctypes.sizeof(_lib.BackvoteSummary),
                     ctypes.sizeof(_lib.Pose), ctypes.sizeof(_lib.VoteParams), ctypes.sizeof(_lib.VoteBuffers), ctypes.sizeof(_lib.InstanceIO),
                     ctypes.sizeof(_lib.Frame), ctypes.sizeof(_lib.ScaleSelect)]


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from cppf2_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.CppfError):
        _lib.load()


def test_sphere_lut_is_a_superset_of_every_possible_hit():
    """Host-built cube-map lookup (cppf_sphere_lut_build): for random and adversarial directions, every lattice
    point passing the reference test dot > cos_thr (eval.py:45) is listed in the direction's cell."""
    import numpy as np
    from cppf2_b200 import _lib
    from cppf2_b200.voting import cos_threshold, fibonacci_sphere
    lib = _lib.load()
    rng = np.random.default_rng(5)
    for S, tol in ((720, 1.0), (360, 2.0), (1440, 0.5)):
        sph = np.ascontiguousarray(np.array(fibonacci_sphere(S), dtype=np.float32))
        thr = cos_threshold(tol)
        built = None
        for g in (32, 48, 64, 96, 128):
            buf = np.empty(int(lib.cppf_sphere_lut_bytes(g)), dtype=np.uint8)
            if lib.cppf_sphere_lut_build(sph.ctypes.data, S, thr, g, buf.ctypes.data) == 0:
                built = (g, buf)
                break
        assert built is not None
        g, buf = built
        hdr = buf[:16].view(np.uint32)
        assert hdr[1] == g and hdr[2] == 4 and hdr[3] == S
        cells = buf[16:].view(np.uint16).reshape(6, g, g, 4)
        p = rng.standard_normal((300000, 3)).astype(np.float32)
        p /= np.linalg.norm(p, axis=-1, keepdims=True)
        # directions on cube edges / corners / cell borders and right on top of lattice points
        edge = np.float32([[1, 1, 0], [1, -1, 0.3], [1, 1, 1], [-1, 1, -1], [0, 0, 1], [0, -1, 0], [1, 0.5, 0.25], [1, 0.0625, -0.125]])
        p = np.concatenate([p, edge / np.linalg.norm(edge, axis=-1, keepdims=True), sph]).astype(np.float32)
        # numpy restatement of the device-side cell assignment (rotation.cu: cube_cell)
        a = np.abs(p)
        axis = np.where((a[:, 0] >= a[:, 1]) & (a[:, 0] >= a[:, 2]), 0, np.where(a[:, 1] >= a[:, 2], 1, 2))
        rows = np.arange(p.shape[0])
        m = p[rows, axis]
        u = p[rows, (axis + 1) % 3] / np.abs(m)
        v = p[rows, (axis + 2) % 3] / np.abs(m)
        iu = np.clip(((u + 1) * np.float32(0.5 * g)).astype(np.int64), 0, g - 1)
        iv = np.clip(((v + 1) * np.float32(0.5 * g)).astype(np.int64), 0, g - 1)
        listed = cells[2 * axis + (m < 0), iv, iu]                      # [rows, 4]
        dots = p @ sph.T                                                  # float32
        hit_r, hit_s = np.nonzero(dots > np.float32(thr))
        assert hit_r.size > 0
        ok = (listed[hit_r] == hit_s[:, None].astype(np.uint16)).any(-1)
        assert ok.all(), f"S={S}: {np.count_nonzero(~ok)} hits missing from the lookup table"
