"""Parity of the HEADLINE configurations themselves (the workloads bench.py times) against the CPU oracle.

The engine runs the full batch (4 096 clips, what one B200 gets); the oracle (`oracle/dragposer_port.py`, the torch restatement of
python/src/drag_pose.py:196-414 pinned against the unmodified reference in tests/test_oracle.py) runs a SAMPLE of those clips --
chosen to cover the first / last CTA, both 16-clip groups of a CTA, both predictor parts (clips below / above 2 048) and the last,
ragged tile -- with exactly the same inputs.  Bars (north_star): joint positions and root within 1 mm on every frame; with early
stopping the per-clip iteration counts within +-1.
"""
import numpy as np
import pytest
import torch

import dragposer_port as port
from dragposer_b200 import synthetic

pytestmark = pytest.mark.gpu

B = 4096
POS_TOL = 1e-3
FIXED = dict(stop_eps_pos=-1.0, stop_eps_rot=-1.0, max_iter=100, min_loss_incr=-float("inf"), learning_rate=1e-2)
EARLY = dict(stop_eps_pos=0.01 * 0.01, stop_eps_rot=0.01, max_iter=100, min_loss_incr=0.00001, learning_rate=1e-2)


def sample_clips(n_random=40, seed=7):
    """Clip indices spread over the launch geometry of 4 096 clips: 147 CTAs x 28 clips (two groups of 14), the last CTA holds the
    8 clips 4088..4095; the predictor runs clips [0, 2048) and [2048, 4096) as two parts on two streams."""
    forced = [0, 1, 13, 14, 15, 27, 28, 29, 41, 42, 2046, 2047, 2048, 2049, 2071, 2072, 4059, 4060, 4087, 4088, 4089, 4094, 4095]
    rng = np.random.default_rng(seed)
    rest = rng.choice(B, n_random + len(forced), replace=False)
    idx = list(dict.fromkeys(forced + [int(i) for i in rest]))[: n_random + len(forced)]
    return np.array(sorted(idx))


def joint_positions(pw, pose_std):
    q = torch.as_tensor(pose_std) * pw.std_q + pw.mean_q
    q = q.reshape(q.shape[0], 22, 4)
    pos, _ = port.fk_chain(port.root_to_local(q, pw.parents), torch.zeros(q.shape[0], 3), pw.offsets, pw.parents)
    return pos.numpy()


def run_both(engine_factory, port_weights, temporal_model, wl, cfg, idx, n_frames, opt, variable):
    eng = engine_factory(B)
    ident = np.tile([[1.0, 0, 0, 0]], (B, 1))
    eng.set_initial_state(wl["latent0"], np.zeros((B, 3)), ident, np.zeros((B, 6)))
    n = len(idx)
    ora = port.PortDragPose(port_weights, temporal_model.sd)
    ora.set_initial_state(wl["latent0"][idx], np.zeros((n, 3)), ident[:n], np.zeros((n, 6)))
    common = dict(lambda_rot=1.0, lambda_temporal=cfg.lambda_temporal, temporal_future_window=cfg.temporal_future_window,
                  joint_adjustment_weight=cfg.joint_adjustment_weight)
    rows = []
    for t in range(n_frames):
        if variable:
            tr = (wl["joints_tb"][t], wl["weights_tb"][t])
            pose, gpos = eng.run(wl["tgt_pos"][t], wl["tgt_rot"][t], *tr, n_ee=wl["n_ee"][t], joint_adjustment_indices=cfg.joint_adjustment,
                                 **common, **opt)
            op, og = ora.run(wl["tgt_pos"][t][idx], wl["tgt_rot"][t][idx], tr[0][idx], tr[1][idx], n_ee=wl["n_ee"][t][idx],
                             joint_adjustment=cfg.joint_adjustment, **common, **opt)
        else:
            pose, gpos = eng.run(wl["tgt_pos"][t], wl["tgt_rot"][t], wl["joints"], wl["weights"], joint_adjustment_indices=cfg.joint_adjustment,
                                 **common, **opt)
            op, og = ora.run(wl["tgt_pos"][t][idx], wl["tgt_rot"][t][idx], wl["joints"], wl["weights"], joint_adjustment=cfg.joint_adjustment,
                             **common, **opt)
        assert eng.last_decoder_path() == 3  # the tcgen05 frame kernel, as in the bench
        iters, _ = eng.frame_stats()
        dpos = np.abs(joint_positions(port_weights, pose[idx]) - joint_positions(port_weights, op.numpy())).max(axis=(1, 2))
        dg = np.abs(gpos[idx] - og.numpy()).max(axis=1)
        rows.append((dpos, dg, iters[idx].copy(), ora.iters.numpy().copy()))
        assert np.isfinite(pose).all() and np.isfinite(gpos).all()
    return rows, eng


def test_headline_6_trackers_4096_clips_100_iterations_vs_oracle(engine_factory, pose_model, model_npz, port_weights, temporal_model):
    """bench.py's default workload: 4 096 clips, 6 trackers, window 0 (predictor every frame), 100 fixed iterations, 4 frames."""
    cfg = synthetic.config_6_trackers()
    T = 4
    wl = synthetic.make_workload(pose_model, model_npz["offsets"], cfg, B, T)
    idx = sample_clips()
    rows, _ = run_both(engine_factory, port_weights, temporal_model, wl, cfg, idx, T, FIXED, variable=False)
    for t, (dpos, dg, it, oit) in enumerate(rows):
        print(f"6 trackers, frame {t}: {len(idx)} sampled clips of {B}: joints max {dpos.max()*1e3:.4f} mm (median {np.median(dpos)*1e3:.5f}), "
              f"root max {dg.max()*1e3:.4f} mm")
        assert (it == 100).all() and (oit == 100).all()
        assert dpos.max() <= POS_TOL and dg.max() <= POS_TOL, (t, idx[dpos.argmax()], dpos.max(), dg.max())


def test_headline_3_trackers_variable_mask_window_16_vs_oracle(engine_factory, pose_model, model_npz, port_weights, temporal_model):
    """BASELINE config 3 as bench.py --trackers 3 runs it: 4 096 clips, head + hands with hands dropping out, window 16, 100 fixed
    iterations, 20 frames -- the predictor runs on frames 0 and 16 (5 autoregressive passes each), so a roll-over of the target
    buffer happens at batch scale."""
    cfg = synthetic.config_3_trackers()
    T = 20
    wl = synthetic.make_workload(pose_model, model_npz["offsets"], cfg, B, T, variable_mask=True)
    # make sure the sample sees both tracker counts (hands drop with probability 0.02 per frame)
    dropped = np.nonzero((wl["n_ee"] < 3).any(axis=0))[0]
    idx = np.array(sorted(set(sample_clips(30).tolist()) | set(dropped[:: max(1, len(dropped) // 24)].tolist())))
    assert (wl["n_ee"][:, idx] == 2).any() and (wl["n_ee"][:, idx] == 3).any()
    rows, eng = run_both(engine_factory, port_weights, temporal_model, wl, cfg, idx, T, FIXED, variable=True)
    worst = 0.0
    for t, (dpos, dg, it, oit) in enumerate(rows):
        worst = max(worst, dpos.max(), dg.max())
        assert (it == 100).all()
        assert dpos.max() <= POS_TOL and dg.max() <= POS_TOL, (t, idx[dpos.argmax()], dpos.max(), dg.max())
    print(f"3 trackers (variable mask), window 16, {T} frames, {len(idx)} sampled clips of {B} ({int((wl['n_ee'][:, idx] == 2).sum())} "
          f"clip-frames with a hand dropped): worst joint / root difference {worst*1e3:.4f} mm")
    st = eng.state(cfg.temporal_future_window)
    assert st["current_index"] == T % 16


def test_headline_3_trackers_early_stop_iteration_counts_vs_oracle(engine_factory, pose_model, model_npz, port_weights, temporal_model):
    """Same batch with the reference's early stopping (eval_drag.py:210-214): per-clip iteration counts within +-1 of the oracle's and
    positions within 1 mm, frame by frame.  A clip whose stop decision slips (a float compare next to its threshold) starts the next
    frame from a different latent; from then on it is a different trajectory, so a clip is compared up to and
    including its first slipped frame and the slips are counted."""
    cfg = synthetic.config_3_trackers()
    T = 20
    wl = synthetic.make_workload(pose_model, model_npz["offsets"], cfg, B, T, variable_mask=True)
    idx = sample_clips(41)
    rows, _ = run_both(engine_factory, port_weights, temporal_model, wl, cfg, idx, T, EARLY, variable=True)
    alive = np.ones(len(idx), bool)
    checked = 0
    for t, (dpos, dg, it, oit) in enumerate(rows):
        assert np.abs(it[alive] - oit[alive]).max() <= 1, (t, it[alive], oit[alive])
        assert dpos[alive].max() <= POS_TOL and dg[alive].max() <= POS_TOL, (t, dpos[alive].max(), dg[alive].max())
        checked += int(alive.sum())
        alive &= it == oit
    mean_it = np.mean([r[2].mean() for r in rows])
    print(f"3 trackers, early stop, {T} frames x {len(idx)} sampled clips: {checked} clip-frames compared, {int((~alive).sum())} clips had a "
          f"+-1 slip of the stop decision, mean {mean_it:.1f} iterations per frame")
    assert alive.mean() >= 0.75
