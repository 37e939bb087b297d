"""GPU: the eval_drag drop-in (BASELINE config #1 on an excerpt) against the reference's own evaluation loop."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_eval_drag_on_bvh_excerpt_matches_reference(tmp_path, monkeypatch):
    from dragposer_b200 import eval_drag, model, motion
    from dragposer_b200.bvh import Bvh

    monkeypatch.chdir(tmp_path)  # the CLI writes data/eval_<name>.bvh relative to the cwd (train.py:505-508)
    g = np.load(os.path.join(G, "ref_eval_bvh.npz"))
    cfg = tmp_path / "6.json"
    from dragposer_b200 import synthetic
    import json

    c = synthetic.config_6_trackers()
    cfg.write_text(json.dumps(dict(mask=c.mask.tolist(), weights=c.weights.tolist(), enable_joint_adjustment=True,
                                   joint_adjustment_indices=[0, 0], joint_adjustment_weight=1.0, lambda_temporal=0.02,
                                   temporal_future_window=0)))
    res = eval_drag.evaluate(os.path.join(G, "model_dancedb.npz"), os.path.join(G, "example_48f.bvh"), str(cfg), quiet=True,
                             initial_latent=g["latent0"])
    assert np.abs(res["iterations"] - g["iters"]).max() <= 1, (res["iterations"], g["iters"])
    pm = model.load_folded_npz(os.path.join(G, "model_dancedb.npz"))
    b = Bvh(os.path.join(G, "example_48f.bvh"))
    par, off = b.skeleton()
    z = np.zeros((48, 3))
    p1, _ = motion.fk_np(motion.result_local_quats(res["poses"], pm.mean_q, pm.std_q, par).astype(np.float64), z, off.astype(np.float64), par)
    p2, _ = motion.fk_np(motion.result_local_quats(g["pose"], pm.mean_q, pm.std_q, par).astype(np.float64), z, off.astype(np.float64), par)
    d = np.abs(p1 - p2).max(axis=(1, 2))
    root = np.abs(res["global_pos"] - g["gpos"]).max()
    rots = b.quaternions()
    m_ref, e_ref = motion.mpjpe(rots, motion.result_local_quats(g["pose"], pm.mean_q, pm.std_q, par).astype(np.float64), off.astype(np.float64), par)
    print(f"eval_drag excerpt: joint diff vs reference, frames 0-7: {d[:8].max()*1e3:.4f} mm, all 48: {d.max()*1e3:.2f} mm; root diff {root*1e3:.4f} mm; "
          f"MPJPE {res['mpjpe']*100:.2f} cm (reference {m_ref*100:.2f}), MPEEPE {res['mpeepe']*100:.2f} cm (reference {e_ref*100:.2f}), "
          f"{48/res['time']:.0f} frames/s")
    # Early-stopped frames run ~2 Adam steps from a FRESH optimiser state, and the first Adam step is lr*sign(g): latent
    # dimensions whose gradient sits in the fp32 noise get a +-lr kick whose sign is not reproducible across
    # implementations, so two fp32-faithful runs drift apart by ~2.5x per frame until they saturate at the ~1.5 cm scale
    # of the method's own error.  The CPU oracle port shows the SAME divergence from the reference on this excerpt
    # (1.6 mm at frame 12, 12-17 mm later).  Reproducible: the first frames to 1 mm, the root
    # (snapped by the joint adjustment), the iteration counts and the accuracy metrics.
    assert d[:8].max() < 1e-3 and root < 1e-3
    assert d.max() < 0.05
    assert abs(res["mpjpe"] - m_ref) < 5e-3 and abs(res["mpeepe"] - e_ref) < 5e-3
    assert res["mpjpe"] < 0.06  # a wrong quaternion convention or FK order gives tens of centimetres
    # the written BVH parses back to the same local rotations
    out = Bvh(res["out_path"])
    q = out.quaternions()
    ref_q = motion.result_local_quats(res["poses"], pm.mean_q, pm.std_q, par)
    dots = np.abs(np.sum(q * ref_q / np.linalg.norm(ref_q, axis=-1, keepdims=True), axis=-1))
    assert dots.min() > 1 - 1e-6
