import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import model, synthetic
from dragposer_b200.engine import BatchedDragPose
npz_path = os.path.join(ROOT, "tests/golden/model_dancedb.npz")
npz = np.load(npz_path); pm = model.load_folded_npz(npz_path)
tm = model.temporal_from_state(model.random_temporal_state(2222))
g = np.load(os.path.join(ROOT, "tests/golden/ref_trace_6trk.npz"))
cfg = synthetic.config_6_trackers()
EARLY = dict(stop_eps_pos=0.01 * 0.01, stop_eps_rot=0.01, max_iter=100, min_loss_incr=0.00001, learning_rate=1e-2)
eng = BatchedDragPose(pm, npz["offsets"], tm, 512)
B = 3
eng.set_initial_state(g["latent0"], np.zeros((B, 3)), np.tile([[1.0, 0, 0, 0]], (B, 1)), np.zeros((B, 6)))
eng.enable_trace(True)
for t in range(3):
    pose, gpos = eng.run(g["tgt_pos"][t], g["tgt_rot"][t], g["joints"], g["weights"], lambda_rot=1, lambda_temporal=cfg.lambda_temporal,
                         temporal_future_window=0, joint_adjustment_indices=cfg.joint_adjustment, joint_adjustment_weight=cfg.joint_adjustment_weight,
                         decoder_path=2, **EARLY)
    iters, losses = eng.frame_stats()
    tr = eng.trace(100)
    st = eng.state(0)
    print("frame", t, "iters", iters, "ref", g["early_iters"][t], "losses", losses.tolist())
    print("  nan in pose per clip", np.isnan(pose).sum(1), "gpos", gpos.tolist())
    for c in range(B):
        n = iters[c]
        print("  clip", c, "trace latent nan iters:", np.where(np.isnan(tr["latent"][c, :n]).any(1))[0][:5], "grad nan:", np.where(np.isnan(tr["grad"][c, :n]).any(1))[0][:5],
              "loss first/last", tr["loss"][c, 0].tolist(), tr["loss"][c, max(n - 1, 0)].tolist())
    print("  state latent nan", np.isnan(st["latent"]).sum(), "grot", st["global_rot"].tolist())
