"""CPU cost of enqueueing one frame (device-resident inputs): how far the host can run ahead of the GPU. Run on a B200."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import model, synthetic
from dragposer_b200.engine import BatchedDragPose, RunOptions
B, T = 4096, 40
npz = os.path.join(ROOT, "tests/golden/model_dancedb.npz")
pm = model.load_folded_npz(npz); off = np.load(npz)["offsets"]
tm = model.temporal_from_state(model.random_temporal_state(2222))
cfg = synthetic.config_6_trackers()
wl = synthetic.make_workload(pm, off, cfg, B, 2)
eng = BatchedDragPose(pm, off, tm, B)
eng.set_initial_state(wl["latent0"], np.zeros((B, 3)), np.tile([[1.0, 0, 0, 0]], (B, 1)), np.zeros((B, 6)))
opts = RunOptions(stop_eps_pos=-1.0, stop_eps_rot=-1.0, max_iter=100, min_loss_incr=-float("inf"), learning_rate=1e-2, lambda_rot=1,
                  lambda_temporal=cfg.lambda_temporal, temporal_future_window=0)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream()
d_tp = torch.from_numpy(wl["tgt_pos"][0]).to(dev); d_tr = torch.from_numpy(wl["tgt_rot"][0]).to(dev)
d_j = torch.from_numpy(wl["joints"].astype(np.int32)).to(dev); d_w = torch.from_numpy(wl["weights"]).to(dev)
d_pose = torch.empty((B, 88), device=dev); d_g = torch.empty((B, 3), device=dev)
for prof in (False, True):
    eng.set_profiling(prof)
    with torch.cuda.stream(st):
        for _ in range(3):
            eng.run_frames_device(1, d_tp, d_tr, d_j, d_w, d_pose, d_g, options=opts, stream=st.cuda_stream)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(T):
            eng.run_frames_device(1, d_tp, d_tr, d_j, d_w, d_pose, d_g, options=opts, stream=st.cuda_stream)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
    print(f"profiling={prof}: enqueue {1e3 * (t1 - t0) / T:.3f} ms/frame, total {1e3 * (t2 - t0) / T:.3f} ms/frame")
