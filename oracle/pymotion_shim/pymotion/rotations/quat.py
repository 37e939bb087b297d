"""Numpy quaternion primitives, (w, x, y, z) order.  ORACLE-ONLY.

Call sites in the reference: train.py:331-335,432-433,482,487,495;
motion_data.py:52-62,254-264; run_drag.py:9,136.
"""
import numpy as np

_AXIS = {"x": 0, "y": 1, "z": 2}


def mul(q0, q1):
    w0, x0, y0, z0 = q0[..., 0], q0[..., 1], q0[..., 2], q0[..., 3]
    w1, x1, y1, z1 = q1[..., 0], q1[..., 1], q1[..., 2], q1[..., 3]
    return np.stack(
        (
            w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1,
            w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
            w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1,
            w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1,
        ),
        axis=-1,
    )


def mul_vec(q, v):
    u = q[..., 1:]
    t = 2.0 * np.cross(u, v)
    return v + q[..., 0:1] * t + np.cross(u, t)


def inverse(q):
    return q * np.array([1.0, -1.0, -1.0, -1.0], dtype=q.dtype)


def length(q):
    return np.sqrt(np.sum(q * q, axis=-1))


def normalize(q, eps=1e-8):
    return q / (length(q)[..., None] + eps)


def _axis_angle(axis_chars, angle):
    """Quaternion for a rotation of `angle` about the per-element axis char."""
    half = 0.5 * angle
    q = np.zeros(angle.shape + (4,), dtype=angle.dtype)
    q[..., 0] = np.cos(half)
    s = np.sin(half)
    for ch, idx in _AXIS.items():
        sel = axis_chars == ch
        q[..., 1 + idx] = np.where(sel, s, 0.0)
    return q


def from_euler(euler, order):
    """q = q(axis0, e0) (x) q(axis1, e1) (x) q(axis2, e2); radians."""
    order = np.asarray(order)
    order = np.broadcast_to(order, euler.shape)
    q0 = _axis_angle(order[..., 0], euler[..., 0])
    q1 = _axis_angle(order[..., 1], euler[..., 1])
    q2 = _axis_angle(order[..., 2], euler[..., 2])
    return mul(q0, mul(q1, q2))


def to_matrix(q):
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    x2, y2, z2 = x + x, y + y, z + z
    xx, yy, wx = x * x2, y * y2, w * x2
    xy, yz, wy = x * y2, y * z2, w * y2
    xz, zz, wz = x * z2, z * z2, w * z2
    m = np.empty(q.shape[:-1] + (3, 3), dtype=q.dtype)
    m[..., 0, 0] = 1.0 - (yy + zz)
    m[..., 0, 1] = xy - wz
    m[..., 0, 2] = xz + wy
    m[..., 1, 0] = xy + wz
    m[..., 1, 1] = 1.0 - (xx + zz)
    m[..., 1, 2] = yz - wx
    m[..., 2, 0] = xz - wy
    m[..., 2, 1] = yz + wx
    m[..., 2, 2] = 1.0 - (xx + yy)
    return m


def to_euler(q, order):
    """Inverse of from_euler for Tait-Bryan orders (three distinct axes)."""
    order = np.asarray(order)
    order = np.broadcast_to(order, q.shape[:-1] + (3,))
    m = to_matrix(q)
    out = np.zeros(q.shape[:-1] + (3,), dtype=q.dtype)
    for row in np.unique(order.reshape(-1, 3), axis=0):
        key = "".join(row)
        i, j, k = (_AXIS[c] for c in key)
        sel = (order[..., 0] == key[0]) & (order[..., 1] == key[1]) & (order[..., 2] == key[2])
        even = (i, j, k) in ((0, 1, 2), (1, 2, 0), (2, 0, 1))
        sgn = 1.0 if even else -1.0
        b = np.arcsin(np.clip(sgn * m[..., i, k], -1.0, 1.0))
        a = np.arctan2(-sgn * m[..., j, k], m[..., k, k])
        c = np.arctan2(-sgn * m[..., i, j], m[..., i, i])
        e = np.stack((a, b, c), axis=-1)
        out = np.where(sel[..., None], e, out)
    return out


def unroll(q, axis=0):
    """Pick the sign of each quaternion that is closest to the previous frame."""
    q = np.swapaxes(q, 0, axis).copy()
    for f in range(1, q.shape[0]):
        d = np.sum(q[f] * q[f - 1], axis=-1)
        q[f][d < 0] *= -1.0
    return np.swapaxes(q, 0, axis)
