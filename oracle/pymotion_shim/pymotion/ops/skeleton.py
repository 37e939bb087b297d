"""Root-space dual quaternion conversion.  ORACLE-ONLY (motion_data.py:58,260)."""
import numpy as np
from ..rotations import quat, dual_quat


def to_root_dual_quat(rotations, global_pos, parents, offsets):
    """rotations (F,J,4) local quats -> (F,J,8) dual quats in ROOT space.

    Joint 0 keeps its world rotation and global_pos; joints whose parent is the
    root keep their local rotation and offset; deeper joints accumulate along
    the chain *excluding* the root's own transform.
    """
    F, J = rotations.shape[0], rotations.shape[1]
    rot = rotations.copy()
    trans = np.zeros((F, J, 3), dtype=rotations.dtype)
    trans[:, 0] = global_pos
    for j in range(1, J):
        p = parents[j]
        if p == 0:
            trans[:, j] = offsets[j]
        else:
            trans[:, j] = trans[:, p] + quat.mul_vec(rot[:, p], np.broadcast_to(offsets[j], (F, 3)))
            rot[:, j] = quat.mul(rot[:, p], rotations[:, j])
    return dual_quat.from_rotation_translation(rot, trans)


def from_root_dual_quat(dq, parents):
    r, t = dual_quat.to_rotation_translation(dq)
    rot = r.copy()
    for j in reversed(range(1, r.shape[-2])):
        p = parents[j]
        if p == 0:
            continue
        rot[..., j, :] = quat.mul(quat.inverse(r[..., p, :]), r[..., j, :])
    return rot, t
