"""Synthetic tracker streams of the shape BASELINE.json names.

Clip c draws from `numpy.random.default_rng(2222 + c)`: the start latent is the
encoder mean of the zero (= mean) pose plus 0.3 N(0,I); every frame the latent
takes a 0.05 N(0,I) random-walk step, is decoded with the (folded) decoder and
run through forward kinematics with the root at the origin; the tracked joints'
positions / rotation matrices are that frame's targets.  Host-side numpy only --
this is data generation, not the hot path.
"""
from __future__ import annotations

import json

import numpy as np

from .model import PoseModel


def quat_to_matrix_np(q):
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    x2, y2, z2 = x + x, y + y, z + z
    m = np.empty(q.shape[:-1] + (3, 3), dtype=q.dtype)
    m[..., 0, 0] = 1.0 - (y * y2 + z * z2)
    m[..., 0, 1] = x * y2 - w * z2
    m[..., 0, 2] = x * z2 + w * y2
    m[..., 1, 0] = x * y2 + w * z2
    m[..., 1, 1] = 1.0 - (x * x2 + z * z2)
    m[..., 1, 2] = y * z2 - w * x2
    m[..., 2, 0] = x * z2 - w * y2
    m[..., 2, 1] = y * z2 + w * x2
    m[..., 2, 2] = 1.0 - (x * x2 + y * y2)
    return m


def pose_fk_np(model: PoseModel, offsets, y, root_rot=None, root_pos=None):
    """Standardised decoder output (B,92) -> joint positions (B,J,3), rotations (B,J,3,3)
    using the closed form R_j = R_0 M(q_j), p_j = p_par + R_par o_j."""
    B = y.shape[0]
    u = (y[:, :88] * model.std_q + model.mean_q).reshape(B, -1, 4)
    q = u / (np.linalg.norm(u, axis=-1, keepdims=True) + np.float32(1e-8))
    M = quat_to_matrix_np(q)
    R0 = M[:, 0] if root_rot is None else root_rot
    R = np.einsum("bij,bnjk->bnik", R0, M)
    R[:, 0] = R0
    J = R.shape[1]
    p = np.zeros((B, J, 3), dtype=y.dtype)
    if root_pos is not None:
        p[:, 0] = root_pos
    for j in range(1, J):
        par = model.parents[j]
        p[:, j] = p[:, par] + np.einsum("bij,j->bi", R[:, par], offsets[j])
    return p, R


class TrackerConfig:
    """The reference's JSON tracker config (`python/config/*.json`, read by
    `python/src/eval_drag.py:33-43`): mask[22], weights[22][2] (pos, rot),
    enable_joint_adjustment, joint_adjustment_indices [joint, ee_slot],
    joint_adjustment_weight, lambda_temporal, temporal_future_window."""

    def __init__(self, data: dict):
        self.mask = np.asarray(data["mask"], dtype=np.float32)
        self.weights = np.asarray(data["weights"], dtype=np.float32)
        self.enable_joint_adjustment = bool(data["enable_joint_adjustment"])
        self.joint_adjustment_indices = tuple(int(i) for i in data["joint_adjustment_indices"])
        self.joint_adjustment_weight = float(data["joint_adjustment_weight"])
        self.lambda_temporal = float(data["lambda_temporal"])
        self.temporal_future_window = int(data["temporal_future_window"])
        if self.mask.shape[0] != self.weights.shape[0] or self.weights.shape[1] != 2:
            raise ValueError("mask must be (J,) and weights (J,2)")

    @classmethod
    def load(cls, path):
        with open(path, "r") as fh:
            return cls(json.load(fh))

    @property
    def joints(self):
        return np.nonzero(self.mask)[0].astype(np.int32)

    @property
    def tracker_weights(self):
        return self.weights[self.joints]

    @property
    def joint_adjustment(self):
        return self.joint_adjustment_indices if self.enable_joint_adjustment else None


_W = [[10, 10], [1, .01], [1, .01], [5, .01], [1, .01], [1, .01], [1, .01], [5, .01], [1, .01], [1, .01], [1, .01],
      [1, .01], [1, .01], [5, .01], [1, .01], [1, .01], [1, .01], [5, .01], [1, .01], [1, .01], [1, .01], [5, .01]]


def _mask(joints):
    return [1 if j in joints else 0 for j in range(22)]


def config_6_trackers():
    """Same content as `python/config/6_trackers_config.json` (= eval_drag.py:68-131 defaults)."""
    return TrackerConfig(dict(mask=_mask((0, 3, 7, 13, 17, 21)), weights=_W, enable_joint_adjustment=True,
                              joint_adjustment_indices=[0, 0], joint_adjustment_weight=1.0,
                              lambda_temporal=0.02, temporal_future_window=0))


def config_3_trackers():
    """Same content as `python/config/3_trackers_config.json` (head weight 20/20)."""
    w = [list(r) for r in _W]
    w[13] = [20, 20]
    return TrackerConfig(dict(mask=_mask((13, 17, 21)), weights=w, enable_joint_adjustment=True,
                              joint_adjustment_indices=[13, 0], joint_adjustment_weight=0.1,
                              lambda_temporal=0.15, temporal_future_window=16))


def make_workload(model: PoseModel, offsets, cfg: TrackerConfig, n_clips, n_frames, first_clip=0,
                  variable_mask=False):
    """Returns dict with latent0 (B,24), tgt_pos (T,B,E,3), tgt_rot (T,B,E,3,3),
    joints (E,), weights (E,2), n_ee (T,B) and, for variable_mask, per-frame
    compacted joints/weights (T,B,E) / (T,B,E,2)."""
    joints, weights = cfg.joints, cfg.tracker_weights
    E = len(joints)
    mu0, _ = model.encode_np(np.zeros((1, 176), np.float32))
    noise = np.empty((n_clips, n_frames + 1, 24), np.float32)
    drop = np.zeros((n_clips, n_frames, E), bool)
    for i in range(n_clips):
        rng = np.random.default_rng(2222 + first_clip + i)
        noise[i] = rng.standard_normal((n_frames + 1, 24), dtype=np.float32)
        if variable_mask:  # slot 0 (head) always on; others drop for U{10..30} frames w.p. 0.02/frame
            start = rng.random((n_frames, E)) < 0.02
            length = rng.integers(10, 31, (n_frames, E))
            for e in range(1, E):
                t = 0
                while t < n_frames:
                    if start[t, e]:
                        drop[i, t : t + length[t, e], e] = True
                        t += int(length[t, e])
                    else:
                        t += 1
    z = mu0 + np.float32(0.3) * noise[:, 0]
    latent0 = z.copy()
    tgt_pos = np.empty((n_frames, n_clips, E, 3), np.float32)
    tgt_rot = np.empty((n_frames, n_clips, E, 3, 3), np.float32)
    for t in range(n_frames):
        z = z + np.float32(0.05) * noise[:, t + 1]
        p, R = pose_fk_np(model, offsets, model.decode_np(z))
        tgt_pos[t], tgt_rot[t] = p[:, joints], R[:, joints]
    out = dict(latent0=latent0, tgt_pos=tgt_pos, tgt_rot=tgt_rot, joints=joints, weights=weights,
               lambda_temporal=cfg.lambda_temporal, window=cfg.temporal_future_window,
               joint_adjustment=cfg.joint_adjustment, joint_adjustment_weight=cfg.joint_adjustment_weight)
    if variable_mask:
        # at least two trackers stay on (the reference cannot run E = 1)
        both = drop[..., 1:].all(-1) if E > 2 else np.zeros(drop.shape[:2], bool)
        drop[both, E - 1] = False
        keep = ~drop
        n_ee = keep.sum(-1).astype(np.int32).T  # (T,B)
        order = np.argsort(~keep, axis=-1, kind="stable")  # kept slots first, original order
        order_t = np.transpose(order, (1, 0, 2))  # (T,B,E)
        out["n_ee"] = n_ee
        out["joints_tb"] = joints[order_t].astype(np.int32)
        out["weights_tb"] = weights[order_t]
        out["tgt_pos"] = np.take_along_axis(tgt_pos, order_t[..., None], 2)
        out["tgt_rot"] = np.take_along_axis(tgt_rot, order_t[..., None, None], 2)
        out["slot_order"] = order_t
        for k in ("n_ee", "joints_tb", "weights_tb", "tgt_pos", "tgt_rot", "slot_order"):  # frame-major, C-contiguous like a recorded stream
            out[k] = np.ascontiguousarray(out[k])
    return out
