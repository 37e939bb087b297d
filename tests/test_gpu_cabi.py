"""GPU: the DragPoserDLL C ABI driven like DragPoserDLL/main.cpp:10-41, against the golden session
recorded from the reference's RunDrag (tests/golden/ref_rundrag.npz)."""
import ctypes as C
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Float3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Quat(C.Structure):
    _fields_ = [("w", C.c_float), ("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Float2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


@pytest.fixture(scope="module")
def dll(tmp_path_factory):
    from dragposer_b200 import build, export_model

    lib = C.CDLL(build.build_all()[1])
    lib.init_drag_poser.restype = C.c_void_p
    V = C.c_void_p
    lib.set_reference_skeleton.argtypes = [V, C.c_char_p]
    lib.load_models.argtypes = [V, C.c_char_p]
    lib.set_mask_and_weights.argtypes = [V, C.POINTER(C.c_float), C.POINTER(Float2)]
    lib.init_drag_model.argtypes = [V, Float3, Quat]
    lib.set_optim_params.argtypes = [V, C.c_float, C.c_float, C.c_int, C.c_float]
    lib.set_lambdas.argtypes = [V, C.c_float, C.c_float, C.c_int]
    lib.set_global_pos.argtypes = [V, Float3]
    lib.drag_pose.argtypes = [V, C.c_int, C.POINTER(Float3), C.POINTER(Quat), C.POINTER(Quat), C.POINTER(Float3)]
    lib.destroy_drag_poser.argtypes = [V]
    lib.dp_last_status.argtypes = [V]
    lib.dp_last_message.argtypes = [V]
    lib.dp_last_message.restype = C.c_char_p
    lib.dp_set_initial_latent.argtypes = [V, C.POINTER(C.c_float)]
    d = tmp_path_factory.mktemp("model")
    export_model.export(os.path.join(ROOT, "tests", "golden", "model_dancedb.npz"), str(d / "model.dpm"), allow_random_temporal=True)
    return lib, str(d)


def open_session(lib, model_dir, g, max_iter, window):
    h = lib.init_drag_poser()
    lib.set_reference_skeleton(h, os.path.join(ROOT, "tests", "golden", "skeleton22.bvh").encode())
    lib.load_models(h, model_dir.encode())
    assert lib.dp_last_status(h) == 0, lib.dp_last_message(h)
    mask = (C.c_float * 22)(*g["mask"].tolist())
    weights = (Float2 * 22)(*[Float2(*w) for w in g["weights"].tolist()])
    lib.set_mask_and_weights(h, mask, weights)
    lib.dp_set_initial_latent(h, g["latent0"].astype(np.float32).ctypes.data_as(C.POINTER(C.c_float)))
    lib.init_drag_model(h, Float3(-2.6648, 0.9977, 3.7518), Quat(0.6381, 0.0078, -0.7698, 0.0110))
    lib.set_optim_params(h, 0.01 * 0.01, 0.01, max_iter, 0.01)
    lib.set_lambdas(h, 1, 0.02, window)
    assert lib.dp_last_status(h) == 0, lib.dp_last_message(h)
    return h


def test_main_cpp_session_matches_reference_rundrag(dll):
    lib, model_dir = dll
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_rundrag.npz"))
    for _ in range(2):  # main.cpp creates and destroys sessions in a loop
        h = open_session(lib, model_dir, g, 10, 60)
        T = g["tgt_pos"].shape[0]
        for t in range(T):
            pos = (Float3 * 6)(*[Float3(*p) for p in g["tgt_pos"][t].tolist()])
            rot = (Quat * 6)(*[Quat(*q) for q in g["tgt_quat"].tolist()])
            res = (Quat * 22)()
            gp = (Float3 * 1)()
            lib.drag_pose(h, 6, pos, rot, res, gp)
            assert lib.dp_last_status(h) == 0, lib.dp_last_message(h)
            q = np.array([[r.w, r.x, r.y, r.z] for r in res], np.float32)
            p = np.array([gp[0].x, gp[0].y, gp[0].z], np.float32)
            # local quaternions within 1e-4 and the root within 1 mm of the reference session
            assert np.abs(q - g["result_pose"][t]).max() < 2e-4, (t, np.abs(q - g["result_pose"][t]).max())
            assert np.abs(p - g["result_gpos"][t, 0]).max() < 1e-3
            lib.set_global_pos(h, gp[0])
        lib.destroy_drag_poser(h)


def test_streaming_latency_and_error_paths(dll):
    lib, model_dir = dll
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_rundrag.npz"))
    h = open_session(lib, model_dir, g, 5, 16)  # Unity VR settings (VRScene.unity:973-979)
    pos = (Float3 * 6)(*[Float3(*p) for p in g["tgt_pos"][0].tolist()])
    rot = (Quat * 6)(*[Quat(*q) for q in g["tgt_quat"].tolist()])
    res, gp = (Quat * 22)(), (Float3 * 1)()
    lat = []
    for i in range(200):
        t0 = time.perf_counter()
        lib.drag_pose(h, 6, pos, rot, res, gp)
        lat.append(time.perf_counter() - t0)
        lib.set_global_pos(h, gp[0])
    assert lib.dp_last_status(h) == 0
    lat = np.array(lat[20:]) * 1e3
    print(f"drag_pose C ABI, B=1, MaxIter 5, window 16: p50 {np.percentile(lat, 50):.3f} ms, p90 {np.percentile(lat, 90):.3f} ms")
    before = np.array([[r.w, r.x, r.y, r.z] for r in res])
    lib.drag_pose(h, 5, pos, rot, res, gp)  # wrong tracker count: reported, outputs untouched, no abort
    assert lib.dp_last_status(h) != 0
    assert np.array_equal(before, np.array([[r.w, r.x, r.y, r.z] for r in res]))
    lib.destroy_drag_poser(h)


def test_two_engines_on_two_devices_in_one_process(pose_model, temporal_model, model_npz):
    """the opt-in to > 48 KB of dynamic shared memory is per device; a second engine on another GPU of the same
    process must run the large-shared-memory kernels too (tensor-core frame kernel, predictor) and give the same results."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from dragposer_b200 import synthetic
    from dragposer_b200.engine import BatchedDragPose

    cfg = synthetic.config_6_trackers()
    B = 600  # >= 512: tcgen05 frame kernel
    wl = synthetic.make_workload(pose_model, model_npz["offsets"], cfg, B, 2)
    outs = []
    engines = [BatchedDragPose(pose_model, model_npz["offsets"], temporal_model, B, device=d) for d in (0, 1)]
    for eng in engines:
        eng.set_initial_state(wl["latent0"], np.zeros((B, 3)), np.tile([[1.0, 0, 0, 0]], (B, 1)), np.zeros((B, 6)))
    for t in range(2):
        outs = [eng.run(wl["tgt_pos"][t], wl["tgt_rot"][t], wl["joints"], wl["weights"], lambda_rot=1, lambda_temporal=cfg.lambda_temporal,
                        temporal_future_window=0, max_iter=10, joint_adjustment_indices=cfg.joint_adjustment,
                        joint_adjustment_weight=cfg.joint_adjustment_weight) for eng in engines]
    assert engines[0].last_decoder_path() == 3 and engines[1].last_decoder_path() == 3
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    for eng in engines:
        eng.close()
