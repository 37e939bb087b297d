"""Timeline of the two groups of CTA 0 of the tcgen05 frame kernel over iterations 40..43 (profiling level 2: the clocked instantiation of the kernel)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import model, synthetic
from dragposer_b200.engine import BatchedDragPose, RunOptions
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
npz = os.path.join(ROOT, "tests/golden/model_dancedb.npz")
pm = model.load_folded_npz(npz); off = np.load(npz)["offsets"]
tm = model.temporal_from_state(model.random_temporal_state(2222))
cfg = synthetic.config_6_trackers()
wl = synthetic.make_workload(pm, off, cfg, B, 3)
eng = BatchedDragPose(pm, off, tm, B)
eng.set_initial_state(wl["latent0"], np.zeros((B, 3), np.float32), np.tile(np.float32([1, 0, 0, 0]), (B, 1)), np.zeros((B, 6), np.float32))
opts = RunOptions(stop_eps_pos=-1.0, stop_eps_rot=-1.0, max_iter=100, min_loss_incr=-float("inf"), learning_rate=1e-2, lambda_rot=1,
                  lambda_temporal=cfg.lambda_temporal, temporal_future_window=0, decoder_path=3)
eng.run(wl["tgt_pos"][0], wl["tgt_rot"][0], wl["joints"], wl["weights"], options=opts)
eng.set_profiling(2)
eng.run(wl["tgt_pos"][1], wl["tgt_rot"][1], wl["joints"], wl["weights"], options=opts)
v = eng.timeline()
t0 = v.min()
names = ["loop top", "fwd done", "kin done", "kin barrier", "bwd2 done", "adam done"]
ev = sorted((int(v[g, i, k] - t0), g, i, k) for g in range(2) for i in range(4) for k in range(6))
for t, g, i, k in ev:
    print(f"{t:7d}  {'            ' * g}g{g} it{40 + i} {names[k]}")
eng.close()
