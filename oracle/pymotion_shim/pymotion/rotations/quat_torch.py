"""Torch quaternion primitives, (w, x, y, z) order.  ORACLE-ONLY.

Call sites in the reference: drag_pose.py:88,102; utils.py:29-30,96;
autoencoder.py:248; loss.py:30-31.
"""
import torch


def mul(q0, q1):
    """Hamilton product q0 (x) q1."""
    w0, x0, y0, z0 = torch.unbind(q0, -1)
    w1, x1, y1, z1 = torch.unbind(q1, -1)
    return torch.stack(
        (
            w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1,
            w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
            w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1,
            w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1,
        ),
        dim=-1,
    )


def mul_vec(q, v):
    """Rotate vector v by quaternion q:  v + w t + u x t,  t = 2 u x v."""
    v = v.to(q.dtype)  # the (commented-out) extension losses of drag_pose.py:140 pass an integer axis vector
    u = q[..., 1:]
    t = 2.0 * torch.cross(u, v, dim=-1)
    return v + q[..., 0:1] * t + torch.cross(u, t, dim=-1)


def inverse(q):
    """Conjugate (inverse of a unit quaternion)."""
    return q * torch.tensor([1.0, -1.0, -1.0, -1.0], dtype=q.dtype, device=q.device)


def length(q):
    return torch.sqrt(torch.sum(q * q, dim=-1))


def normalize(q, eps=1e-8):
    return q / (length(q).unsqueeze(-1) + eps)


def from_matrix(m):
    """Rotation matrix (..., 3, 3) -> quaternion (w, x, y, z); largest-component (Shepperd) selection, differentiable.
    Only the commented-out extension losses of drag_pose.py:141,145 use it (re-enabled by oracle/reference_harness.py)."""
    m00, m01, m02 = m[..., 0, 0], m[..., 0, 1], m[..., 0, 2]
    m10, m11, m12 = m[..., 1, 0], m[..., 1, 1], m[..., 1, 2]
    m20, m21, m22 = m[..., 2, 0], m[..., 2, 1], m[..., 2, 2]
    tr = m00 + m11 + m22
    cand = torch.stack((tr, m00, m11, m22), -1)
    which = torch.argmax(cand, -1)

    def safe(x):
        return torch.sqrt(torch.clamp(x, min=1e-12))

    s0 = safe(1.0 + tr) * 2.0
    q0 = torch.stack((0.25 * s0, (m21 - m12) / s0, (m02 - m20) / s0, (m10 - m01) / s0), -1)
    s1 = safe(1.0 + m00 - m11 - m22) * 2.0
    q1 = torch.stack(((m21 - m12) / s1, 0.25 * s1, (m01 + m10) / s1, (m02 + m20) / s1), -1)
    s2 = safe(1.0 + m11 - m00 - m22) * 2.0
    q2 = torch.stack(((m02 - m20) / s2, (m01 + m10) / s2, 0.25 * s2, (m12 + m21) / s2), -1)
    s3 = safe(1.0 + m22 - m00 - m11) * 2.0
    q3 = torch.stack(((m10 - m01) / s3, (m02 + m20) / s3, (m12 + m21) / s3, 0.25 * s3), -1)
    allq = torch.stack((q0, q1, q2, q3), -2)
    idx = which[..., None, None].expand(which.shape + (1, 4))
    return torch.gather(allq, -2, idx).squeeze(-2)
