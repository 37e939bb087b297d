import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from dragposer_b200 import model, synthetic
from dragposer_b200.engine import BatchedDragPose
npz_path = os.path.join(ROOT, "tests/golden/model_dancedb.npz")
npz = np.load(npz_path); pm = model.load_folded_npz(npz_path)
tm = model.temporal_from_state(model.random_temporal_state(2222))
cfg = synthetic.config_6_trackers()
B, T = 64, 3
wl = synthetic.make_workload(pm, npz["offsets"], cfg, B, T)
engs = {p: BatchedDragPose(pm, npz["offsets"], tm, 64) for p in (1, 2)}
for p, e in engs.items():
    e.set_initial_state(wl["latent0"], np.zeros((B, 3)), np.tile([[1., 0, 0, 0]], (B, 1)), np.zeros((B, 6)))
    e.enable_trace(True)
for t in range(T):
    tr = {}
    for p, e in engs.items():
        e.run(wl["tgt_pos"][t], wl["tgt_rot"][t], wl["joints"], wl["weights"], stop_eps_pos=-1.0, stop_eps_rot=-1.0, max_iter=100,
              min_loss_incr=-float("inf"), learning_rate=1e-2, lambda_rot=1, lambda_temporal=cfg.lambda_temporal, temporal_future_window=0,
              joint_adjustment_indices=cfg.joint_adjustment, joint_adjustment_weight=cfg.joint_adjustment_weight, decoder_path=p)
        tr[p] = e.trace(100)
    dz = np.abs(tr[1]["latent"] - tr[2]["latent"]).max(axis=(0, 2))
    dg = np.abs(tr[1]["grad"] - tr[2]["grad"]).max(axis=(0, 2))
    gn = np.abs(tr[1]["grad"]).max(axis=(0, 2))
    print(f"frame {t}: latent diff @iter 0,1,5,20,50,99: {dz[[0,1,5,20,50,99]]}")
    print(f"          grad diff: {dg[[0,1,5,20,50,99]]}  |grad| {gn[[0,1,5,20,50,99]]}")
    s1, s2 = engs[1].state(0), engs[2].state(0)
    for k in ("latent", "global_pos", "global_rot", "latent_buf", "disp_buf", "height_buf", "target_buf"):
        print(f"          state {k}: max diff {np.abs(s1[k]-s2[k]).max():.3e}")
