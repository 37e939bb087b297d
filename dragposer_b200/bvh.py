"""Small BVH reader (hierarchy + motion) for the boundary classes.

`RunDrag.set_reference_skeleton` (python/src/run_drag.py:30-38) and `eval_drag.py:45-48`
get parents / offsets / rotations from a BVH through `train.get_info_from_bvh`
(python/src/train.py:322-341): quaternions from the per-joint euler channels, parents[0]
forced to 0 and the root offset forced to zero.  End sites are not joints.
"""
from __future__ import annotations

import numpy as np

from . import rotations as rot

_AXIS = {"x": 1, "y": 2, "z": 3}


class Bvh:
    def __init__(self, path):
        with open(path, "r") as fh:
            tok = fh.read().split()
        self.names, self.parents, self.offsets, self.channels, self.end_sites = [], [], [], [], []
        stack, i, in_end, cur = [], 0, False, -1
        while tok[i] != "MOTION":
            t = tok[i]
            if t in ("ROOT", "JOINT"):
                self.names.append(tok[i + 1])
                self.parents.append(stack[-1] if stack else 0)
                self.offsets.append([0.0, 0.0, 0.0])
                self.channels.append([])
                cur = len(self.names) - 1
                i += 2
            elif t == "End":
                in_end = True
                i += 2
            elif t == "{":
                stack.append(-1 if in_end else cur)
                i += 1
            elif t == "}":
                if stack.pop() == -1:
                    in_end = False
                cur = stack[-1] if stack else -1
                i += 1
            elif t == "OFFSET":
                off = [float(tok[i + 1]), float(tok[i + 2]), float(tok[i + 3])]
                if in_end:
                    self.end_sites.append((stack[-2], off))
                else:
                    self.offsets[cur] = off
                i += 4
            elif t == "CHANNELS":
                n = int(tok[i + 1])
                self.channels[cur] = tok[i + 2 : i + 2 + n]
                i += 2 + n
            else:
                i += 1
        self.n_frames = int(tok[i + 2])
        self.frame_time = float(tok[i + 5])
        n_ch = sum(len(c) for c in self.channels)
        vals = np.array(tok[i + 6 : i + 6 + self.n_frames * n_ch], dtype=np.float64).reshape(self.n_frames, n_ch)
        J = len(self.names)
        self.raw_offsets = np.array(self.offsets, dtype=np.float64)
        self.positions = np.tile(self.raw_offsets[None], (self.n_frames, 1, 1))
        self.euler = np.zeros((self.n_frames, J, 3))
        self.rot_order = [[] for _ in range(J)]
        col = 0
        for j in range(J):
            r = 0
            for ch in self.channels[j]:
                axis = ch[0].lower()
                if ch.lower().endswith("position"):
                    self.positions[:, j, "xyz".index(axis)] = vals[:, col]
                else:
                    self.euler[:, j, r] = vals[:, col]
                    self.rot_order[j].append(axis)
                    r += 1
                col += 1

    def skeleton(self):
        """(parents list with parents[0] = 0, offsets (J,3) float32 with a zero root offset)."""
        off = self.raw_offsets.astype(np.float32).copy()
        off[0] = 0.0
        par = [int(p) for p in self.parents]
        par[0] = 0
        return par, off

    def quaternions(self):
        """(F,J,4) unit quaternions: q = q(axis0,e0) * q(axis1,e1) * q(axis2,e2), sign-unrolled over time."""
        F, J = self.euler.shape[:2]
        half = np.radians(self.euler) * 0.5
        q = np.zeros((F, J, 4))
        q[..., 0] = 1.0
        for j in range(J):
            qj = None
            for r, axis in enumerate(self.rot_order[j]):
                qa = np.zeros((F, 4))
                qa[:, 0] = np.cos(half[:, j, r])
                qa[:, _AXIS[axis]] = np.sin(half[:, j, r])
                qj = qa if qj is None else rot.mul(qj, qa)
            if qj is not None:
                q[:, j] = qj
        for f in range(1, F):
            flip = np.sum(q[f] * q[f - 1], axis=-1) < 0
            q[f][flip] *= -1.0
        return rot.normalize(q)
