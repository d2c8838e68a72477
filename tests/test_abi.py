"""CPU: the C-ABI library builds/loads and exports every symbol include/butd_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "butd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bd_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_nine_reference_ops():
    syms = declared_symbols()
    for s in ("bd_fps", "bd_gather_points", "bd_gather_points_grad", "bd_ball_query", "bd_group_points",
              "bd_group_points_grad", "bd_three_nn", "bd_three_interpolate", "bd_three_interpolate_grad"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from butd_detr_b200 import build
    lib = ctypes.CDLL(build.build())
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in butd_b200.h but not exported"
    lib.bd_arch.restype = ctypes.c_char_p
    assert lib.bd_arch() == b"sm_100a"
    assert lib.bd_version() >= 100


def test_python_binding_covers_the_header():
    from butd_detr_b200 import _lib
    assert set(declared_symbols()) <= set(_lib.EXPORTED) | {"bd_version", "bd_last_error", "bd_arch"}


def test_invalid_arguments_return_error_codes_not_exit():
    """Argument validation happens before any CUDA call, so it is testable without a GPU."""
    from butd_detr_b200 import _lib
    lib = _lib.load()
    assert lib.bd_fps(None, 3, 1, 16, 4, None, None, None) == 1
    assert b"null" in lib.bd_last_error()
    assert lib.bd_ball_query(None, None, 3, 1, 1, 1, 0.2, 4, None, None) == 1
    assert lib.bd_linear_f32(None, 0, None, 0, None, None, None, 0, 1, 1, 1, 0, None) == 1


def test_tuning_hooks_validate_their_argument():
    """Host-only switches: bad values are refused with an error code, good ones accepted (and restored)."""
    from butd_detr_b200 import _lib
    lib = _lib.load()
    for w in (8, 16, 32, 0):  # 8 = two CTAs per SM, 0 = chosen by the number of scenes (default, restored last)
        assert lib.bd_fps_grid_set_warps(w) == 0
    assert lib.bd_fps_grid_set_warps(12) == 1
    assert b"bd_fps_grid_set_warps" in lib.bd_last_error()
    assert lib.bd_fps_set_cluster(5) == 1
    assert lib.bd_fps_set_cluster(-1) == 0


def test_sass_uses_cluster_and_async_instructions():
    """The FPS kernel must really be the cluster/DSMEM design (st.async + mbarrier)."""
    import subprocess
    from butd_detr_b200 import build
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    assert "REDUX" in sass   # warp reductions (CREDUX.MAX on sm_100)
    assert "STAS" in sass    # st.async into peer CTAs' shared memory
    assert "SYNCS" in sass   # mbarrier arrive / try_wait


def test_sass_uses_blackwell_tensor_and_tma_instructions():
    """The tensor-core kernels must really be tcgen05 / TMEM / TMA code (B200_PROFILING.md: the PTX
    names never appear in SASS): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG =
    cp.async.bulk.tensor (activations of the linear kernel), UBLKCP = cp.async.bulk (weights, K/V
    tiles, write-out), and no legacy mma.sync (HMMA) anywhere."""
    import subprocess
    from butd_detr_b200 import build
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR"):
        assert mnemonic in sass, mnemonic
    assert " HMMA" not in sass
    # griddepcontrol.wait / launch_dependents of the programmatic dependent launches
    assert "ACQBULK" in sass or "PREEXIT" in sass or "DEPBAR" in sass
