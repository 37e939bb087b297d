"""Debug aid for tests/test_gpu_headline.py (3-tracker variable-mask workload): per-frame joint difference of chosen clips between the
tcgen05 kernel (path 3), the fp32 CUDA-core kernel (path 1) and the oracle port, plus the oracle's own spread under 1e-7..1e-5
perturbations of the start latent.  Tells a conditioning problem (both kernels drift away from the oracle and from each other at the
same pace) from a defect (one kernel jumps).  Usage: debug_headline3.py clip [clip ...]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dragposer_port as port
from dragposer_b200 import model, synthetic
from dragposer_b200.engine import BatchedDragPose
import test_gpu_headline as H

clips = np.array([int(c) for c in sys.argv[1:]] or [2108])
npz_path = os.path.join(ROOT, "tests/golden/model_dancedb.npz")
npz = np.load(npz_path); pm = model.load_folded_npz(npz_path)
tm = model.temporal_from_state(model.random_temporal_state(2222)); pw = port.PortWeights(npz)
cfg = synthetic.config_3_trackers(); T = 20; B = H.B
wl = synthetic.make_workload(pm, npz["offsets"], cfg, B, T, variable_mask=True)
ident = np.tile([[1.0, 0, 0, 0]], (B, 1))
common = dict(lambda_rot=1.0, lambda_temporal=cfg.lambda_temporal, temporal_future_window=cfg.temporal_future_window,
              joint_adjustment_weight=cfg.joint_adjustment_weight)
out = {}
for path in (3, 1):
    eng = BatchedDragPose(pm, npz["offsets"], tm, B)
    eng.set_initial_state(wl["latent0"], np.zeros((B, 3)), ident, np.zeros((B, 6)))
    res = []
    for t in range(T):
        pose, gpos = eng.run(wl["tgt_pos"][t], wl["tgt_rot"][t], wl["joints_tb"][t], wl["weights_tb"][t], n_ee=wl["n_ee"][t],
                             joint_adjustment_indices=cfg.joint_adjustment, decoder_path=path, **common, **H.FIXED)
        res.append((H.joint_positions(pw, pose[clips]), gpos[clips].copy()))
    out[path] = res
    eng.close()
n = len(clips); NP = H.N_PERT
rows_idx = np.tile(clips, 1 + NP); m = len(rows_idx)
lat = wl["latent0"][rows_idx].copy(); rng = np.random.default_rng(123)
for k in range(NP):
    lat[(k + 1) * n:(k + 2) * n] += rng.normal(0, (1e-7, 1e-6, 1e-5)[k * 3 // NP], (n, 24)).astype(np.float32)
ora = port.PortDragPose(pw, tm.sd)
ora.set_initial_state(lat, np.zeros((m, 3)), ident[:m], np.zeros((m, 6)))
print("frame n_ee | path3-oracle  path1-oracle  path3-path1  oracle spread   [mm, max over joints and root], per clip", clips.tolist())
for t in range(T):
    op, og = ora.run(wl["tgt_pos"][t][rows_idx], wl["tgt_rot"][t][rows_idx], wl["joints_tb"][t][rows_idx], wl["weights_tb"][t][rows_idx],
                     n_ee=wl["n_ee"][t][rows_idx], joint_adjustment=cfg.joint_adjustment, **common, **H.FIXED)
    opos = H.joint_positions(pw, op.numpy()).reshape(1 + NP, n, 22, 3); ogp = og.numpy().reshape(1 + NP, n, 3)
    f = lambda a, b, ga, gb: np.maximum(np.abs(a - b).max(axis=(1, 2)), np.abs(ga - gb).max(axis=1)) * 1e3
    d3 = f(out[3][t][0], opos[0], out[3][t][1], ogp[0]); d1 = f(out[1][t][0], opos[0], out[1][t][1], ogp[0])
    d31 = f(out[3][t][0], out[1][t][0], out[3][t][1], out[1][t][1])
    sp = np.maximum(np.abs(opos[1:] - opos[:1]).max(axis=(0, 2, 3)), np.abs(ogp[1:] - ogp[:1]).max(axis=(0, 2))) * 1e3
    print(t, wl["n_ee"][t][clips].tolist(), "|", np.round(d3, 4).tolist(), np.round(d1, 4).tolist(), np.round(d31, 4).tolist(), np.round(sp, 4).tolist())
