"""Torch forward kinematics.  ORACLE-ONLY (eval_drag.py:190, eval_metrics.py:14,24)."""
import torch
from ..rotations import quat_torch


def _to_matrix4(q):
    w, x, y, z = torch.unbind(q, -1)
    x2, y2, z2 = x + x, y + y, z + z
    xx, yy, wx = x * x2, y * y2, w * x2
    xy, yz, wy = x * y2, y * z2, w * y2
    xz, zz, wz = x * z2, z * z2, w * z2
    m = torch.zeros(q.shape[:-1] + (4, 4), dtype=q.dtype, device=q.device)
    m[..., 0, 0] = 1.0 - (yy + zz)
    m[..., 0, 1] = xy - wz
    m[..., 0, 2] = xz + wy
    m[..., 1, 0] = xy + wz
    m[..., 1, 1] = 1.0 - (xx + zz)
    m[..., 1, 2] = yz - wx
    m[..., 2, 0] = xz - wy
    m[..., 2, 1] = yz + wx
    m[..., 2, 2] = 1.0 - (xx + yy)
    m[..., 3, 3] = 1.0
    return m


def fk(rot, global_pos, offsets, parents):
    """rot (...,J,4) local quats, global_pos (...,3) -> positions (...,J,3), rotmats (...,J,3,3)."""
    m = _to_matrix4(quat_torch.normalize(rot))
    m[..., :3, 3] = offsets
    m[..., 0, :3, 3] = global_pos
    out = [m[..., 0, :, :]]
    for i in range(1, len(parents)):
        out.append(torch.matmul(out[int(parents[i])], m[..., i, :, :]))
    m = torch.stack(out, dim=-3)
    return m[..., :3, 3], m[..., :3, :3]
