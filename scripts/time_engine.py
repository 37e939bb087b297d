import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import model
from dragposer_b200.engine import BatchedDragPose
npz = os.path.join(ROOT, "tests/golden/model_dancedb.npz")
pm = model.load_folded_npz(npz); off = np.load(npz)["offsets"]
t = time.time(); tm = model.temporal_from_state(model.random_temporal_state(2222)); print("temporal state", time.time() - t)
for B in (4, 256):
    t = time.time(); eng = BatchedDragPose(pm, off, tm, 512); print("engine create", time.time() - t)
    eng.set_initial_state(np.zeros((B, 24)), np.zeros((B, 3)), np.tile([[1., 0, 0, 0]], (B, 1)), np.zeros((B, 6)))
    for path in (0, 1, 0):
        eng.set_predictor_path(path)
        for W in (0, 16):
            t = time.time(); eng.predict_targets(W); print(f"B={B} path={path} W={W} predict", time.time() - t)
    eng.close()
