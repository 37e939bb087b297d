import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import model, synthetic
from dragposer_b200.engine import BatchedDragPose
npz = os.path.join(ROOT, "tests/golden/model_dancedb.npz")
pm = model.load_folded_npz(npz); off = np.load(npz)["offsets"]
tm = model.temporal_from_state(model.random_temporal_state(2222))
cfg = synthetic.config_3_trackers()
B, T = 2100, 2
wl = synthetic.make_workload(pm, off, cfg, B, T, variable_mask=True)
perm = np.random.default_rng(1).permutation(B)
for path in (1, 3):
    for W in (16, 0):
        kw = dict(lambda_rot=1, lambda_temporal=cfg.lambda_temporal, temporal_future_window=W, max_iter=15,
                  joint_adjustment_indices=cfg.joint_adjustment, joint_adjustment_weight=cfg.joint_adjustment_weight, decoder_path=path)
        outs = []
        for order in (np.arange(B), perm):
            eng = BatchedDragPose(pm, off, tm, B)
            eng.set_initial_state(wl["latent0"][order], np.zeros((B, 3)), np.tile([[1.0, 0, 0, 0]], (B, 1)), np.zeros((B, 6)))
            tb = None
            for t in range(T):
                res = eng.run(wl["tgt_pos"][t][order], wl["tgt_rot"][t][order], wl["joints_tb"][t][order], wl["weights_tb"][t][order], n_ee=wl["n_ee"][t][order], **kw)
                if t == 0: tb = eng.state(W)["target_buf"].copy()
            it, ls = eng.frame_stats()
            outs.append((res[0], res[1], it, ls, tb))
            eng.close()
        names = ("pose", "gpos", "iters", "losses", "target_buf")
        msg = []
        for nm, a, b in zip(names, outs[0], outs[1]):
            d = np.abs(a[perm].astype(np.float64) - b.astype(np.float64)).reshape(B, -1).max(1)
            bad = np.nonzero(d > 0)[0]
            msg.append(f"{nm}: {len(bad)} clips differ (max {d.max():.3g}; first permuted rows {bad[:6].tolist()} = clips {perm[bad[:6]].tolist()})")
        print(f"path {path} window {W}:", "; ".join(msg))
