"""Host-side (numpy) clip preparation and result export for the evaluation CLI.

Restates, for ONE clip at a time and outside the hot path:
  * `TestMotionData.add_motion/normalize` (python/src/motion_data.py:225-324): BVH quaternions -> root-space dual
    quaternions (22 x 8) with the incremental root rotation and root-space displacement in the root slot, heights of
    [0,4,8,13,17,21], standardisation with the model statistics;
  * the per-frame target construction of python/src/eval_drag.py:164-202;
  * `result_to_bvh` (python/src/train.py:437-509) and `eval_pos_error` (python/src/eval_metrics.py:6-32).
The `upc-pymotion==0.1.10` helpers these call are restated in `rotations.py` / here.
"""
from __future__ import annotations

import numpy as np

from . import rotations as rot

_AXIS = {"x": 0, "y": 1, "z": 2}


def dq_from_rotation_translation(r, t):
    tq = np.concatenate((np.zeros(t.shape[:-1] + (1,), t.dtype), t), axis=-1)
    return np.concatenate((r, 0.5 * rot.mul(tq, r)), axis=-1)


def dq_to_rotation_translation(dq):
    r = dq[..., :4]
    return r, (2.0 * rot.mul(dq[..., 4:], rot.inverse(r)))[..., 1:]


def to_root_dual_quat(rotations, root_pos, parents, offsets):
    """Local quats (F,J,4) -> (F,J,8): root keeps its world rotation + root_pos; other joints accumulate rotation and
    translation from the root's children down (the root's own transform is excluded)."""
    F, J = rotations.shape[:2]
    r = rotations.copy()
    t = np.zeros((F, J, 3), rotations.dtype)
    t[:, 0] = root_pos
    for j in range(1, J):
        p = parents[j]
        if p == 0:
            t[:, j] = offsets[j]
        else:
            t[:, j] = t[:, p] + rot.mul_vec(r[:, p], np.broadcast_to(offsets[j], (F, 3)))
            r[:, j] = rot.mul(r[:, p], rotations[:, j])
    return dq_from_rotation_translation(r, t)


def fk_np(local_q, root_pos, offsets, parents):
    """pymotion `fk`: local quats (...,J,4), root position (...,3) -> positions (...,J,3), rotmats (...,J,3,3)."""
    R = rot.to_matrix(rot.normalize(local_q))
    J = R.shape[-3]
    pos = np.zeros(R.shape[:-2] + (3,), R.dtype)
    out = R.copy()
    pos[..., 0, :] = root_pos
    for j in range(1, J):
        p = parents[j]
        out[..., j, :, :] = out[..., p, :, :] @ R[..., j, :, :]
        pos[..., j, :] = pos[..., p, :] + np.einsum("...ij,j->...i", out[..., p, :, :], offsets[j])
    return pos, out


def to_euler(q, order):
    """Inverse of Bvh.quaternions' composition q = q(a0,e0) q(a1,e1) q(a2,e2) for one Tait-Bryan axis order (radians)."""
    i, j, k = (_AXIS[c] for c in order)
    m = rot.to_matrix(q)
    sgn = 1.0 if (i, j, k) in ((0, 1, 2), (1, 2, 0), (2, 0, 1)) else -1.0
    b = np.arcsin(np.clip(sgn * m[..., i, k], -1.0, 1.0))
    a = np.arctan2(-sgn * m[..., j, k], m[..., k, k])
    c = np.arctan2(-sgn * m[..., i, j], m[..., i, i])
    return np.stack((a, b, c), axis=-1)


class ClipData:
    """Standardised evaluation inputs of one BVH clip (TestMotionData of the reference, window_size 1)."""

    def __init__(self, rotations, root_positions, parents, offsets, mean_dqs, std_dqs, height_joints=(0, 4, 8, 13, 17, 21)):
        rotations = np.asarray(rotations, np.float64)
        gpos = np.asarray(root_positions, np.float32)
        F = rotations.shape[0]
        disp = np.concatenate((np.zeros((1, 3), np.float32), gpos[1:] - gpos[:-1]), 0)
        disp = rot.mul_vec(rot.inverse(rotations[:, 0]), disp.astype(np.float64))
        incr = rotations[:, 0].copy()
        incr[1:] = rot.mul(rot.inverse(rotations[:-1, 0]), rotations[1:, 0])
        incr[0] = (1.0, 0.0, 0.0, 0.0)
        dqs = to_root_dual_quat(rotations, np.zeros((F, 3)), parents, np.asarray(offsets, np.float64))
        r, t = dq_to_rotation_translation(dqs)
        world_t = rot.mul_vec(r[:, 0:1], t) + gpos[:, None, :].astype(np.float64)
        self.heights = world_t[:, list(height_joints), 1].astype(np.float32)
        dqs[:, 0, :4] = incr
        for f in range(1, F):  # dual-quaternion unroll: pick the cover closest to the previous frame
            flip = np.sum(dqs[f, :, :4] * dqs[f - 1, :, :4], axis=-1) < 0
            dqs[f][flip] *= -1.0
        dqs[:, 0, 4:7] = disp
        dqs[:, 0, 7] = 0.0
        self.dqs = ((dqs.reshape(F, -1).astype(np.float32) - mean_dqs) / std_dqs).astype(np.float32)  # (F,176) standardised
        self.global_pos = gpos                                  # (F,3)
        self.global_rot = rotations[:, 0].astype(np.float32)    # (F,4)
        self.n_frames = F


def frame_targets(clip: ClipData, i, mean_q, std_q, parents, offsets, current_global_pos, joints):
    """eval_drag.py:164-202: ground-truth frame i -> tracker positions (E,3) relative to the current root position and
    world rotation matrices (E,3,3)."""
    q = clip.dqs[i].reshape(-1, 8)[:, :4].reshape(-1) * std_q + mean_q
    q = q.reshape(-1, 4).astype(np.float32)
    q[0] = clip.global_rot[i]
    local = rot.from_root_quat(q[None], parents)[0]
    pos, R = fk_np(local, clip.global_pos[i] - current_global_pos, offsets, parents)
    return pos[joints].astype(np.float32), R[joints].astype(np.float32)


def world_targets(clip: ClipData, mean_q, std_q, parents, offsets, joints, n_frames=None):
    """All frames at once: tracker positions (F,E,3) in WORLD coordinates (no dependence on the optimiser's root estimate) and
    world rotation matrices (F,E,3,3).  With dp_run_params.targets_world the engine subtracts its current global position
    itself, which is what eval_drag.py:164-202 does on the host once per frame."""
    F = clip.n_frames if n_frames is None else n_frames
    q = clip.dqs[:F].reshape(F, -1, 8)[:, :, :4] * np.asarray(std_q).reshape(1, -1, 4) + np.asarray(mean_q).reshape(1, -1, 4)
    q = q.astype(np.float32)
    q[:, 0] = clip.global_rot[:F]
    local = rot.from_root_quat(q, parents)
    pos, R = fk_np(local, clip.global_pos[:F], offsets, parents)
    return pos[:, joints].astype(np.float32), R[:, joints].astype(np.float32)


def result_local_quats(results_pose, mean_q, std_q, parents):
    """(F,88) standardised result -> (F,J,4) parent-local quaternions (train.py:466-484 with are_root_rot_incr=False)."""
    q = (results_pose * std_q + mean_q).reshape(results_pose.shape[0], -1, 4)
    return rot.from_root_quat(q, parents)


def mpjpe(gt_local, res_local, offsets, parents, sparse_joints=(4, 8, 13, 17, 21)):
    """eval_metrics.py:6-32 with the root at the origin in both skeletons."""
    z = np.zeros(gt_local.shape[:1] + (3,))
    pg, _ = fk_np(gt_local, z, offsets, parents)
    pr, _ = fk_np(res_local, z, offsets, parents)
    err = np.linalg.norm(pr - pg[: pr.shape[0]], axis=-1)
    return float(err.mean()), float(err[:, list(sparse_joints)].mean())
