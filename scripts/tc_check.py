import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dragposer_port as port
from dragposer_b200 import model, synthetic
from dragposer_b200.engine import BatchedDragPose
npz_path = os.path.join(ROOT, "tests/golden/model_dancedb.npz")
npz = np.load(npz_path); pm = model.load_folded_npz(npz_path)
tm = model.temporal_from_state(model.random_temporal_state(2222))
cfg = synthetic.config_6_trackers()
n = 100
rng = np.random.default_rng(11)
wl = synthetic.make_workload(pm, npz["offsets"], cfg, n, 1)
lat = wl["latent0"] + 0.2 * rng.standard_normal((n, 24)).astype(np.float32)
grot = rng.standard_normal((n, 4)).astype(np.float32); grot /= np.linalg.norm(grot, axis=1, keepdims=True)
tl = rng.standard_normal((n, 24)).astype(np.float32) * 0.3
eng = BatchedDragPose(pm, npz["offsets"], tm, 4096)
pw64 = port.PortWeights(npz, dtype=torch.float64)
t64 = port.loss_and_grad(pw64, lat, grot, wl["tgt_pos"][0], wl["tgt_rot"][0], tl, wl["joints"], wl["weights"], lambda_rot=1.0, lambda_temporal=0.02, dtype=torch.float64)
for path in (1, 2, 3):
    r = eng.eval_gradient(lat, grot, tl, wl["tgt_pos"][0], wl["tgt_rot"][0], wl["joints"], wl["weights"], lambda_rot=1.0, lambda_temporal=0.02, decoder_path=path)
    rel = np.linalg.norm(r["grad"] - t64["grad"], axis=1) / np.linalg.norm(t64["grad"], axis=1)
    print(f"path {path}: grad rel err max {rel.max():.2e} median {np.median(rel):.2e}; pos err {np.abs(r['pos'] - t64['pos']).max():.2e}; lp err {np.abs(r['lp']-t64['lp']).max():.2e}")
# frames: 64 clips x 3 frames, TC vs SIMT
B, T = 64, 3
wl = synthetic.make_workload(pm, npz["offsets"], cfg, B, T)
outs = {}
for path in (1, 2, 3):
    eng.set_initial_state(wl["latent0"], np.zeros((B, 3)), np.tile([[1., 0, 0, 0]], (B, 1)), np.zeros((B, 6)))
    res = []
    for t in range(T):
        t0 = time.time()
        res.append(eng.run(wl["tgt_pos"][t], wl["tgt_rot"][t], wl["joints"], wl["weights"], stop_eps_pos=-1.0, stop_eps_rot=-1.0, max_iter=100,
                           min_loss_incr=-float("inf"), learning_rate=1e-2, lambda_rot=1, lambda_temporal=cfg.lambda_temporal, temporal_future_window=0,
                           joint_adjustment_indices=cfg.joint_adjustment, joint_adjustment_weight=cfg.joint_adjustment_weight, decoder_path=path))
    outs[path] = res
    print("path", path, "ran; last_decoder_path", eng.last_decoder_path(), "frame time", time.time() - t0)
pw = port.PortWeights(npz)
def positions(pose):
    q = torch.as_tensor(pose) * pw.std_q + pw.mean_q
    q = q.reshape(q.shape[0], 22, 4)
    pos, _ = port.fk_chain(port.root_to_local(q, pw.parents), torch.zeros(q.shape[0], 3), pw.offsets, pw.parents)
    return pos.numpy()
for t in range(T):
    for pth in (2, 3):
        d = np.abs(positions(outs[1][t][0]) - positions(outs[pth][t][0])).max()
        print(f"frame {t}: path {pth} vs fp32 max joint diff {d*1e3:.4f} mm, root diff {np.abs(outs[1][t][1]-outs[pth][t][1]).max()*1e3:.4f} mm")
