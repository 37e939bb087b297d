"""ORACLE-ONLY: writes tests/golden/skeleton22.bvh (hierarchy + one zero frame) from the skeleton stored in
tests/golden/model_dancedb.npz (parents / offsets of python/data/example/eval/example.bvh:2-122)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
z = np.load(os.path.join(ROOT, "tests", "golden", "model_dancedb.npz"))
par, off = z["parents"], z["offsets"]
names = ["Hips", "LeftUpLeg", "LeftLeg", "LeftFoot", "LeftToe", "RightUpLeg", "RightLeg", "RightFoot", "RightToe", "Spine", "Chest",
         "UpperChest", "Neck", "Head", "LeftCollar", "LeftShoulder", "LeftElbow", "LeftWrist", "RightCollar", "RightShoulder",
         "RightElbow", "RightWrist"]  # DragPoserUnity/Assets/Scripts/Core/DragPoser.cs:277-301
children = {j: [c for c in range(1, 22) if par[c] == j] for j in range(22)}
out = ["HIERARCHY"]


def emit(j, d):
    t = "\t" * d
    out.append(f"{t}{'ROOT' if j == 0 else 'JOINT'} {names[j]}")
    out.append(t + "{")
    out.append(f"{t}\tOFFSET {off[j, 0]:.6f} {off[j, 1]:.6f} {off[j, 2]:.6f}")
    out.append(f"{t}\t" + ("CHANNELS 6 Xposition Yposition Zposition Xrotation Yrotation Zrotation" if j == 0
                          else "CHANNELS 3 Xrotation Yrotation Zrotation"))
    for c in children[j]:
        emit(c, d + 1)
    if not children[j]:
        out.extend([f"{t}\tEnd Site", t + "\t{", f"{t}\t\tOFFSET 0.000000 0.000000 0.000000", t + "\t}"])
    out.append(t + "}")


emit(0, 0)
out += ["MOTION", "Frames: 1", "Frame Time: 0.008333", " ".join(["0.000000"] * (6 + 21 * 3))]
with open(os.path.join(ROOT, "tests", "golden", "skeleton22.bvh"), "w") as fh:
    fh.write("\n".join(out) + "\n")
