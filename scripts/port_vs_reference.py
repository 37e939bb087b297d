"""Per-core speed of the oracle port (oracle/dragposer_port.py) against the UNMODIFIED reference loop (python/src/drag_pose.py under
the pymotion shim), same clip, same targets, fixed 100 iterations, one thread.  Build container only (needs /root/reference);
the result is quoted by bench.py (`cpu_baseline.port_vs_unmodified_reference_per_core`) and DESIGN.md."""
import os, sys, time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dragposer_port as port
import reference_harness as rh
from dragposer_b200 import model as dpm, synthetic

torch.set_num_threads(1)
ref = rh.Reference()
ref.temporal.load_state_dict(dpm.random_temporal_state(2222)); ref.temporal.eval()
npz_path = os.path.join(ROOT, "tests/golden/model_dancedb.npz")
npz = np.load(npz_path); pm = dpm.load_folded_npz(npz_path)
cfg = synthetic.config_6_trackers()
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4
wl = synthetic.make_workload(pm, npz["offsets"], cfg, 1, n_frames)
fixed = dict(stop_eps_pos=-1.0, stop_eps_rot=-1.0, max_iter=100, min_loss_incr=-float("inf"), learning_rate=1e-2)

drag = ref.new_drag()
drag.set_initial_pose(torch.zeros(1, 176, 1), torch.zeros(1, 3, 1), torch.tensor([[1.0, 0, 0, 0]]).unsqueeze(-1), torch.zeros(6))
z = torch.from_numpy(wl["latent0"].copy())
drag.latent = z.clone().requires_grad_(); drag.latent_buffer = torch.tile(z, (60, 1))
t_ref = []
for t in range(n_frames):
    t0 = time.perf_counter()
    drag.run(torch.from_numpy(wl["tgt_pos"][t, 0].copy()), torch.from_numpy(wl["tgt_rot"][t, 0].copy()), torch.from_numpy(wl["joints"]).long(),
             torch.from_numpy(wl["weights"].copy()), ref.offsets, lambda_rot=1, lambda_temporal=cfg.lambda_temporal, temporal_future_window=0,
             joint_adjustment_indices=cfg.joint_adjustment, joint_adjustment_weight=cfg.joint_adjustment_weight, **fixed)
    t_ref.append(time.perf_counter() - t0)
ora = port.PortDragPose(port.PortWeights(npz), dpm.random_temporal_state(2222))
ora.set_initial_state(wl["latent0"], np.zeros((1, 3)), [[1.0, 0, 0, 0]], np.zeros((1, 6)))
t_port = []
for t in range(n_frames):
    t0 = time.perf_counter()
    ora.run(wl["tgt_pos"][t], wl["tgt_rot"][t], wl["joints"], wl["weights"], lambda_rot=1.0, lambda_temporal=cfg.lambda_temporal,
            temporal_future_window=0, joint_adjustment=cfg.joint_adjustment, joint_adjustment_weight=cfg.joint_adjustment_weight, **fixed)
    t_port.append(time.perf_counter() - t0)
r, p = np.mean(t_ref[1:]), np.mean(t_port[1:])
print(f"reference {1/r:.2f} frames/s/core ({r*1e3:.0f} ms per frame), port {1/p:.2f} frames/s/core ({p*1e3:.0f} ms per frame), ratio {r/p:.2f} "
      f"({n_frames - 1} timed frames of 100 fixed iterations each, 1 thread)")
