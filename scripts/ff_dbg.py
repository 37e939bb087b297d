import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import model
from dragposer_b200.engine import BatchedDragPose
npz = os.path.join(ROOT, "tests/golden/model_dancedb.npz")
pm = model.load_folded_npz(npz); off = np.load(npz)["offsets"]
tm = model.temporal_from_state(model.random_temporal_state(2222))
B = 4096
eng = BatchedDragPose(pm, off, tm, B)
eng.set_initial_state(np.zeros((B, 24)), np.zeros((B, 3)), np.tile([[1., 0, 0, 0]], (B, 1)), np.zeros((B, 6)))
for i in range(3): eng.predict_targets(0)
torch.cuda.synchronize()
t = time.time()
for i in range(10): eng.lib.dp_engine_predict_targets(eng.h, 0, None)
torch.cuda.synchronize()
print("DP_FF_DBG", os.environ.get("DP_FF_DBG"), "predictor ms", (time.time() - t) / 10 * 1e3)
