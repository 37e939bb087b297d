"""Host-side numpy quaternion helpers ((w,x,y,z) order) used by the boundary classes.

They restate the few `upc-pymotion==0.1.10` functions the reference's boundary code calls
(`python/src/run_drag.py:9,136`, `python/src/train.py:409-434`); the hot path itself never
runs on the host.
"""
from __future__ import annotations

import numpy as np


def mul(a, b):
    w0, x0, y0, z0 = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    w1, x1, y1, z1 = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack((w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1, w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                     w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1, w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1), axis=-1)


def inverse(q):
    return q * np.array([1.0, -1.0, -1.0, -1.0], dtype=q.dtype)


def mul_vec(q, v):
    u = q[..., 1:]
    t = 2.0 * np.cross(u, v)
    return v + q[..., 0:1] * t + np.cross(u, t)


def normalize(q, eps=1e-8):
    return q / (np.sqrt(np.sum(q * q, axis=-1, keepdims=True)) + eps)


def to_matrix(q):
    """Quaternions (...,4) -> rotation matrices (...,3,3), the `1 - 2(yy+zz)` form of utils.py:34-76."""
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


def from_root_quat(q, parents):
    """Root-space quaternions (...,J,4) -> parent-local quaternions (train.py:409-434):
    joints whose parent is the root stay; deeper joints become inverse(q_parent) * q_j."""
    out = q.copy()
    for j in reversed(range(1, q.shape[-2])):
        p = parents[j]
        if p == 0:
            continue
        out[..., j, :] = mul(inverse(out[..., p, :]), out[..., j, :])
    return out
