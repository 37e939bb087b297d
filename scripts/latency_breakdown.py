import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from dragposer_b200 import model, synthetic
from dragposer_b200.engine import BatchedDragPose
npz = "/root/repo/tests/golden/model_dancedb.npz"
pm = model.load_folded_npz(npz); off = np.load(npz)["offsets"]
tm = model.temporal_from_state(model.random_temporal_state(2222))
cfg = synthetic.config_6_trackers()
wl = synthetic.make_workload(pm, off, cfg, 1, 64)
eng = BatchedDragPose(pm, off, tm, 1)
eng.set_initial_state(wl["latent0"], np.zeros((1, 3)), [[1.0, 0, 0, 0]], np.zeros((1, 6)))
for mi in (5, 100):
    kw = dict(stop_eps_pos=1e-4, stop_eps_rot=1e-2, max_iter=mi, min_loss_incr=1e-5, learning_rate=1e-2, lambda_rot=1, lambda_temporal=0.02, temporal_future_window=16)
    for t in range(16): eng.run(wl["tgt_pos"][t], wl["tgt_rot"][t], wl["joints"], wl["weights"], **kw)
    eng.set_profiling(True)
    ts = []
    for t in range(16, 64):
        t0 = time.perf_counter(); eng.run(wl["tgt_pos"][t], wl["tgt_rot"][t], wl["joints"], wl["weights"], **kw); ts.append(time.perf_counter() - t0)
    it, _ = eng.frame_stats()
    mp, mf, n = eng.profile()
    print(f"max_iter {mi}: frame kernel {1e3*mf/n:.1f} us avg, predictor {1e3*mp/n:.1f} us avg per frame, python call p50 {1e6*np.median(ts):.1f} us, last iters {it}")
    eng.set_profiling(False)
