"""Writes the native model file the C-ABI library loads without Python in the loop.

    python -m dragposer_b200.export_model <model_dir | folded.npz> [out.dpm]

    python -m dragposer_b200.export_model <model_dir | folded.npz> [out.dpm] [--random-temporal]

`<model_dir>` is the reference layout (generator.pt, data.pt, temporal.pt,
python/src/train.py:285-319).  A missing temporal.pt is an error, as in the reference
(train_temporal.load_model); `--random-temporal` writes the seed-2222 random-init
predictor instead and records that in the header.  Layout of the .dpm file: b"DPM1",
uint32 version (2), uint32 n_floats, uint32 flags (bit 0: the predictor was read from a
temporal.pt), then float32 arrays in the order of `DPM_FIELDS` followed by the temporal
blob (csrc/dp_temporal.cuh order) and means/stds of the latent.  The DLL refuses a
model.dpm that is older than a temporal.pt lying next to it (re-export).
"""
from __future__ import annotations

import os
import struct
import sys

import numpy as np

from . import model as dpm
from .engine import pack_temporal

DPM_FIELDS = (("A0", 40 * 24), ("b0", 40), ("A1", 60 * 40), ("b1", 60), ("A2", 92 * 60), ("b2", 92), ("mean_q", 88), ("std_q", 88),
              ("mean_d", 3), ("std_d", 3), ("encA0", 112 * 176), ("encb0", 112), ("encA1", 72 * 112), ("encb1", 72),
              ("encA2", 48 * 72), ("encb2", 48), ("mu_w", 24 * 48), ("mu_b", 24), ("lv_w", 24 * 48), ("lv_b", 24))


DPM_FLAG_TRAINED_TEMPORAL = 1


def export(model_dir, out_path=None, parents=dpm.DEFAULT_PARENTS, allow_random_temporal=False):
    pm = dpm.load_pose_model(model_dir, parents)
    tdir = model_dir if os.path.isdir(model_dir) else os.path.dirname(model_dir)
    tm = dpm.load_temporal_model(tdir, allow_random=allow_random_temporal)
    arrs = dict(A0=pm.A[0], b0=pm.b[0], A1=pm.A[1], b1=pm.b[1], A2=pm.A[2], b2=pm.b[2], mean_q=pm.mean_q, std_q=pm.std_q,
                mean_d=pm.mean_d, std_d=pm.std_d, encA0=pm.enc_A[0], encb0=pm.enc_b[0], encA1=pm.enc_A[1], encb1=pm.enc_b[1],
                encA2=pm.enc_A[2], encb2=pm.enc_b[2], mu_w=pm.enc_mu[0], mu_b=pm.enc_mu[1], lv_w=pm.enc_logvar[0],
                lv_b=pm.enc_logvar[1])
    parts = []
    for name, n in DPM_FIELDS:
        a = np.ascontiguousarray(arrs[name], dtype=np.float32).reshape(-1)
        assert a.size == n, (name, a.size, n)
        parts.append(a)
    parts += [pack_temporal(tm), np.asarray(tm.means_latent, np.float32), np.asarray(tm.stds_latent, np.float32)]
    flat = np.concatenate(parts)
    if out_path is None:
        out_path = os.path.join(tdir, "model.dpm")
    with open(out_path, "wb") as fh:
        fh.write(b"DPM1" + struct.pack("<III", 2, flat.size, DPM_FLAG_TRAINED_TEMPORAL if tm.trained else 0))
        fh.write(flat.tobytes())
    return out_path


if __name__ == "__main__":
    argv = [a for a in sys.argv[1:] if a != "--random-temporal"]
    print(export(argv[0], argv[1] if len(argv) > 1 else None, allow_random_temporal="--random-temporal" in sys.argv))
