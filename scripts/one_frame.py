"""A few frames of the headline workload (4096 clips, 6 trackers, 100 fixed iterations) -- the target of `ncu -k regex:dp_frame`."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import model, synthetic
from dragposer_b200.engine import BatchedDragPose, RunOptions

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
trk = sys.argv[3] if len(sys.argv) > 3 else "6"
npz = os.path.join(ROOT, "tests/golden/model_dancedb.npz")
pm = model.load_folded_npz(npz); off = np.load(npz)["offsets"]
tm = model.temporal_from_state(model.random_temporal_state(2222))
cfg = synthetic.config_3_trackers() if trk == "3" else synthetic.config_6_trackers()
wl = synthetic.make_workload(pm, off, cfg, B, T, variable_mask=(trk == "3"))
eng = BatchedDragPose(pm, off, tm, B)
eng.set_initial_state(wl["latent0"], np.zeros((B, 3), np.float32), np.tile(np.float32([1, 0, 0, 0]), (B, 1)), np.zeros((B, 6), np.float32))
opts = RunOptions(stop_eps_pos=-1.0, stop_eps_rot=-1.0, max_iter=100, min_loss_incr=-float("inf"), learning_rate=1e-2, lambda_rot=1,
                  lambda_temporal=cfg.lambda_temporal, temporal_future_window=cfg.temporal_future_window,
                  joint_adjustment_indices=cfg.joint_adjustment, joint_adjustment_weight=cfg.joint_adjustment_weight)
eng.set_profiling(1)
for t in range(T):
    if trk == "3":
        eng.run(wl["tgt_pos"][t], wl["tgt_rot"][t], wl["joints_tb"][t], wl["weights_tb"][t], n_ee=wl["n_ee"][t], options=opts)
    else:
        eng.run(wl["tgt_pos"][t], wl["tgt_rot"][t], wl["joints"], wl["weights"], options=opts)
ms = eng.profile()
print(f"{B} clips, {trk} trackers: predictor {ms[0] / ms[2]:.3f} ms, frame kernel {ms[1] / ms[2]:.3f} ms per frame over {ms[2]} frames")
eng.close()
