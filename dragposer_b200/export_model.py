"""Writes the native model file the C-ABI library loads without Python in the loop.

    python -m dragposer_b200.export_model <model_dir | folded.npz> [out.dpm]

`<model_dir>` is the reference layout (generator.pt, data.pt[, temporal.pt],
python/src/train.py:285-319).  Layout of the .dpm file: b"DPM1", uint32 version,
uint32 n_floats, then float32 arrays in the order of `DPM_FIELDS` followed by the
temporal blob (csrc/dp_temporal.cuh order) and means/stds of the latent.
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


def export(model_dir, out_path=None, parents=dpm.DEFAULT_PARENTS):
    pm = dpm.load_pose_model(model_dir, parents)
    tdir = model_dir if os.path.isdir(model_dir) else os.path.dirname(model_dir)
    tm = dpm.load_temporal_model(tdir)
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
        fh.write(b"DPM1" + struct.pack("<II", 1, flat.size))
        fh.write(flat.tobytes())
    return out_path


if __name__ == "__main__":
    print(export(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None))
