import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import eval_drag, model, motion
from dragposer_b200.bvh import Bvh
G = os.path.join(ROOT, "tests/golden")
g = np.load(os.path.join(G, "ref_eval_bvh.npz"))
res = eval_drag.evaluate(os.path.join(G, "model_dancedb.npz"), os.path.join(G, "example_48f.bvh"), None, quiet=True, initial_latent=g["latent0"], save=False)
pm = model.load_folded_npz(os.path.join(G, "model_dancedb.npz"))
b = Bvh(os.path.join(G, "example_48f.bvh")); par, off = b.skeleton()
z = np.zeros((48, 3))
p1, _ = motion.fk_np(motion.result_local_quats(res["poses"], pm.mean_q, pm.std_q, par).astype(np.float64), z, off.astype(np.float64), par)
p2, _ = motion.fk_np(motion.result_local_quats(g["pose"], pm.mean_q, pm.std_q, par).astype(np.float64), z, off.astype(np.float64), par)
d = np.abs(p1 - p2).max(axis=(1, 2))
for i in range(48):
    print(i, "iters", res["iterations"][i], g["iters"][i], "joint diff mm %.4f" % (d[i] * 1e3), "root diff mm %.4f" % (np.abs(res["global_pos"][i] - g["gpos"][i]).max() * 1e3))
