"""Minimal BVH reader/writer.  ORACLE-ONLY (train.py:322-341,476-508).

`data` keys: names, offsets (J,3), parents (root None), rot_order (J,3) chars,
positions (F,J,3), rotations (F,J,3) degrees in channel order, frame_time,
end_sites (list of (parent, offset)).
"""
import numpy as np
from ..rotations import quat


class BVH:
    def __init__(self):
        self.data = None

    def load(self, path):
        with open(path, "r") as f:
            tokens = f.read().split()
        names, offsets, parents, channels, end_sites = [], [], [], [], []
        stack = []
        i = 0
        in_end = False
        while tokens[i] != "MOTION":
            t = tokens[i]
            if t in ("ROOT", "JOINT"):
                names.append(tokens[i + 1])
                parents.append(stack[-1] if stack else None)
                offsets.append(None)
                channels.append([])
                cur = len(names) - 1
                i += 2
            elif t == "End":
                in_end = True
                i += 2
            elif t == "{":
                stack.append(-1 if in_end else cur)
                i += 1
            elif t == "}":
                popped = stack.pop()
                if popped == -1:
                    in_end = False
                if stack:
                    cur = stack[-1]
                i += 1
            elif t == "OFFSET":
                off = [float(tokens[i + 1]), float(tokens[i + 2]), float(tokens[i + 3])]
                if in_end:
                    end_sites.append((stack[-2], off))
                else:
                    offsets[cur] = off
                i += 4
            elif t == "CHANNELS":
                n = int(tokens[i + 1])
                channels[cur] = tokens[i + 2 : i + 2 + n]
                i += 2 + n
            else:
                i += 1
        assert tokens[i + 1] == "Frames:"
        n_frames = int(tokens[i + 2])
        assert tokens[i + 3] == "Frame" and tokens[i + 4] == "Time:"
        frame_time = float(tokens[i + 5])
        values = np.array(tokens[i + 6 :], dtype=np.float64)
        n_ch = sum(len(c) for c in channels)
        values = values.reshape(n_frames, n_ch)
        J = len(names)
        offsets = np.array(offsets, dtype=np.float64)
        positions = np.tile(offsets[None], (n_frames, 1, 1))
        rotations = np.zeros((n_frames, J, 3))
        rot_order = np.empty((J, 3), dtype="<U1")
        col = 0
        for j in range(J):
            r = 0
            for ch in channels[j]:
                axis = ch[0].lower()
                if ch.endswith("position"):
                    positions[:, j, "xyz".index(axis)] = values[:, col]
                else:
                    rotations[:, j, r] = values[:, col]
                    rot_order[j, r] = axis
                    r += 1
                col += 1
        self.data = {
            "names": names,
            "offsets": offsets,
            "parents": parents,
            "rot_order": rot_order,
            "positions": positions,
            "rotations": rotations,
            "frame_time": frame_time,
            "end_sites": end_sites,
            "channels": channels,
        }
        return self.data

    def get_data(self):
        d = self.data
        order = np.tile(d["rot_order"], (d["rotations"].shape[0], 1, 1))
        rots = quat.unroll(quat.from_euler(np.radians(d["rotations"]), order=order), axis=0)
        parents = list(d["parents"])
        parents[0] = 0
        return rots, d["positions"], parents, d["offsets"], d["end_sites"], None

    def save(self, path):
        d = self.data
        names, parents = d["names"], d["parents"]
        children = {j: [] for j in range(len(names))}
        for j, p in enumerate(parents):
            if p is not None and j != 0:
                children[p].append(j)
        ends = {}
        for p, off in d["end_sites"]:
            ends.setdefault(p, []).append(off)
        lines = ["HIERARCHY"]

        def emit(j, depth):
            tab = "\t" * depth
            lines.append(f"{tab}{'ROOT' if j == 0 else 'JOINT'} {names[j]}")
            lines.append(tab + "{")
            o = d["offsets"][j]
            lines.append(f"{tab}\tOFFSET {o[0]:.6f} {o[1]:.6f} {o[2]:.6f}")
            lines.append(f"{tab}\tCHANNELS {len(d['channels'][j])} " + " ".join(d["channels"][j]))
            for c in children[j]:
                emit(c, depth + 1)
            for off in ends.get(j, []):
                lines.append(f"{tab}\tEnd Site")
                lines.append(tab + "\t{")
                lines.append(f"{tab}\t\tOFFSET {off[0]:.6f} {off[1]:.6f} {off[2]:.6f}")
                lines.append(tab + "\t}")
            lines.append(tab + "}")

        emit(0, 0)
        n_frames = d["rotations"].shape[0]
        lines.append("MOTION")
        lines.append(f"Frames: {n_frames}")
        lines.append(f"Frame Time: {d['frame_time']:.6f}")
        rows = []
        for f in range(n_frames):
            vals = []
            for j in range(len(names)):
                r = 0
                for ch in d["channels"][j]:
                    if ch.endswith("position"):
                        vals.append(d["positions"][f, j, "xyz".index(ch[0].lower())])
                    else:
                        vals.append(d["rotations"][f, j, r])
                        r += 1
            rows.append(" ".join(f"{v:.6f}" for v in vals))
        with open(path, "w") as fh:
            fh.write("\n".join(lines) + "\n" + "\n".join(rows) + "\n")
