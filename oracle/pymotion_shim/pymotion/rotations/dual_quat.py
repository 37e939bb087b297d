"""Dual quaternion helpers, dq = [q_r, 0.5 (0,t) (x) q_r].  ORACLE-ONLY.

Call sites: motion_data.py:61,68,263,272,359-360.
"""
import numpy as np
from . import quat


def from_rotation_translation(r, t):
    tq = np.concatenate((np.zeros(t.shape[:-1] + (1,), dtype=t.dtype), t), axis=-1)
    d = 0.5 * quat.mul(tq, r)
    return np.concatenate((r, d), axis=-1)


def to_rotation_translation(dq):
    r = dq[..., :4]
    d = dq[..., 4:]
    t = 2.0 * quat.mul(d, quat.inverse(r))
    return r, t[..., 1:]


def unroll(dq, axis=0):
    dq = np.swapaxes(dq, 0, axis).copy()
    for f in range(1, dq.shape[0]):
        d = np.sum(dq[f][..., :4] * dq[f - 1][..., :4], axis=-1)
        dq[f][d < 0] *= -1.0
    return np.swapaxes(dq, 0, axis)
