"""Phase clock of the tcgen05 frame kernel (CTA 0): cycles per iteration in each phase. Run on a B200."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import model, synthetic
from dragposer_b200.engine import BatchedDragPose, RunOptions

B, T, ITERS = (int(sys.argv[1]) if len(sys.argv) > 1 else 4096), 6, 100
npz = os.path.join(ROOT, "tests/golden/model_dancedb.npz")
pm = model.load_folded_npz(npz); off = np.load(npz)["offsets"]
tm = model.temporal_from_state(model.random_temporal_state(2222))
cfg = synthetic.config_6_trackers()
wl = synthetic.make_workload(pm, off, cfg, B, T)
for path in (3,):
    eng = BatchedDragPose(pm, off, tm, B)
    eng.set_initial_state(wl["latent0"], np.zeros((B, 3), np.float32), np.tile(np.float32([1, 0, 0, 0]), (B, 1)), np.zeros((B, 6), np.float32))
    opts = RunOptions(stop_eps_pos=-1.0, stop_eps_rot=-1.0, max_iter=ITERS, min_loss_incr=-float("inf"), learning_rate=1e-2, lambda_rot=1,
                      lambda_temporal=cfg.lambda_temporal, temporal_future_window=0, decoder_path=path)
    eng.run(wl["tgt_pos"][0], wl["tgt_rot"][0], wl["joints"], wl["weights"], options=opts)
    eng.set_profiling(2)
    n = 0
    for t in range(1, T):
        eng.run(wl["tgt_pos"][t], wl["tgt_rot"][t], wl["joints"], wl["weights"], options=opts); n += 1
    cyc = eng.phase_cycles(); ms = eng.profile()
    per = [c / (n * ITERS) for c in cyc]
    print(f"path {path}: cycles/iteration forward {per[0]:.0f}  kinematics {per[4]:.0f} (+{per[1]:.0f} scaling/barrier)  backward {per[2]:.0f}  adam {per[3]:.0f}  total {sum(per):.0f};"
          f" frame kernel {ms[1] / ms[2]:.3f} ms")
    eng.close()
