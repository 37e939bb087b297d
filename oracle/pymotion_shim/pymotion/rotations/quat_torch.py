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
