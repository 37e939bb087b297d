"""CPU checks of the C-ABI boundary: both shared libraries load without a GPU and export
every symbol that include/*.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([a-z_][a-z0-9_]*)\s*\([^;{]*\)\s*;", text)
    return sorted(set(n for n in names if n not in ("defined",)))


@pytest.fixture(scope="module")
def built():
    from dragposer_b200 import build

    return build.build_all()


def test_engine_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built[0])
    names = declared_symbols("dp_engine.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    from dragposer_b200 import _lib

    assert sorted(_lib.SIGNATURES) == names  # the ctypes binding covers exactly the header
    lib.dp_engine_version.restype = ctypes.c_int
    assert lib.dp_engine_version() >= 100
    lib.dp_engine_temporal_blob_floats.restype = ctypes.c_size_t
    assert lib.dp_engine_temporal_blob_floats() == 1283976


def test_dragposer_dll_exports_reference_abi(built):
    lib = ctypes.CDLL(built[1])
    reference_abi = ["init_drag_poser", "set_reference_skeleton", "load_models", "set_mask_and_weights", "init_drag_model",
                     "set_optim_params", "set_lambdas", "set_global_pos", "drag_pose", "destroy_drag_poser"]  # exportFunc.h:61-70
    for n in reference_abi + declared_symbols("exportFunc.h"):
        assert hasattr(lib, n), n


def test_dll_session_errors_do_not_cross_the_boundary(built, tmp_path):
    """Host-only behaviour of the shim: bad inputs set a status instead of throwing/aborting."""
    lib = ctypes.CDLL(built[1])
    lib.init_drag_poser.restype = ctypes.c_void_p
    lib.dp_last_status.argtypes = [ctypes.c_void_p]
    lib.dp_last_message.argtypes = [ctypes.c_void_p]
    lib.dp_last_message.restype = ctypes.c_char_p
    lib.set_reference_skeleton.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    lib.dp_get_num_joints.argtypes = [ctypes.c_void_p]
    lib.destroy_drag_poser.argtypes = [ctypes.c_void_p]
    h = lib.init_drag_poser()
    assert h
    lib.set_reference_skeleton(h, str(tmp_path / "missing.bvh").encode())
    assert lib.dp_last_status(h) != 0 and b"cannot open" in lib.dp_last_message(h)
    lib.set_reference_skeleton(h, os.path.join(ROOT, "tests", "golden", "skeleton22.bvh").encode())
    assert lib.dp_last_status(h) == 0 and lib.dp_get_num_joints(h) == 22
    lib.set_mask_and_weights.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    mask = (ctypes.c_float * 22)(*([0.0] * 21 + [1.0]))
    w = (ctypes.c_float * 44)(*([1.0] * 44))
    lib.set_mask_and_weights(h, mask, w)
    assert lib.dp_last_status(h) != 0  # a single tracker is rejected (the reference cannot run E = 1)
    lib.destroy_drag_poser(h)
