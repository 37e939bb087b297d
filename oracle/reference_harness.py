"""ORACLE-ONLY test infrastructure -- NOT part of the product path.

Imports the UNMODIFIED DragPoser reference modules from `/root/reference/python/src`
(available only in the build container; the GPU box has no /root/reference) on top
of the `pymotion` shim in `oracle/pymotion_shim`.  Used by `oracle/make_golden.py`
to generate the committed fixtures under `tests/golden/` and by the CPU tests that
validate `oracle/dragposer_port.py` against the real reference when it is present.

Nothing in `dragposer_b200/` imports this file.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("DRAGPOSER_REFERENCE", "/root/reference")
REFERENCE_SRC = os.path.join(REFERENCE_ROOT, "python", "src")
MODEL_DIR = os.path.join(REFERENCE_ROOT, "python", "models", "model_dancedb")
EXAMPLE_BVH = os.path.join(REFERENCE_ROOT, "python", "data", "example", "eval", "example.bvh")
CONFIG_DIR = os.path.join(REFERENCE_ROOT, "python", "config")
SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pymotion_shim")


def available():
    return os.path.isdir(REFERENCE_SRC)


def activate():
    """Put the shim and the reference sources on sys.path (idempotent)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_SRC)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    for p in (REFERENCE_SRC, SHIM_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)


def load_drag_pose_with_extension_losses():
    """The reference's `DragPose` with the commented-out "Additional Losses" block of drag_pose.py:129-183 (feet-floor, head-hips
    forward, head-hips colinear, hips-feet colinear -- the documented constraints-as-losses extension point) re-enabled: the source
    text is read from the reference tree, the comment markers of exactly that block are stripped, `additional_losses = 0` is dropped
    and the result is exec'd into a fresh module.  Nothing is written to disk and no reference source enters this repository."""
    activate()
    import types

    with open(os.path.join(REFERENCE_SRC, "drag_pose.py"), "r") as fh:
        lines = fh.read().split("\n")
    a = next(i for i, l in enumerate(lines) if l.strip().startswith("# Additional Losses"))
    b = next(i for i, l in enumerate(lines) if l.strip() == "additional_losses = 0")
    out = lines[: a + 1]
    for l in lines[a + 1 : b]:
        ind = len(l) - len(l.lstrip())
        body = l.lstrip()
        if body.startswith("# "):
            body = body[2:]
        elif body == "#":
            body = ""
        out.append(" " * ind + body)
    out += lines[b + 1 :]
    mod = types.ModuleType("drag_pose_extension_losses")
    mod.__file__ = os.path.join(REFERENCE_SRC, "drag_pose.py")
    exec(compile("\n".join(out), mod.__file__ + " [extension losses enabled]", "exec"), mod.__dict__)
    return mod.DragPose


class Reference:
    """Builds the reference objects exactly like eval_drag.main / RunDrag do
    (eval_drag.py:21-59, run_drag.py:16-59) with the two documented deviations
    forced by the missing `temporal.pt` blob: the predictor keeps its seed-2222
    random initialisation and means_latent = 0, stds_latent = 1."""

    def __init__(self, seed=2222, quiet=True):
        activate()
        import contextlib
        import io
        import random

        import numpy as np
        import torch

        import train
        import train_temporal
        from generator_architecture import Generator_Model
        from temporal_transformer import Temporal
        from train_data import Train_Data

        self.torch = torch
        self.train = train
        self.train_temporal = train_temporal
        torch.manual_seed(seed)
        random.seed(seed)
        np.random.seed(seed)
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext():
            bvh = train.get_bvh_from_disk(os.path.dirname(EXAMPLE_BVH), os.path.basename(EXAMPLE_BVH))
            self.rots, self.pos, self.parents, self.offsets_np, self.bvh = train.get_info_from_bvh(bvh)
            self.train_data = Train_Data("cpu", train.param, None)
            self.generator = Generator_Model("cpu", train.param, self.parents, self.train_data).to("cpu")
            self.temporal = Temporal(train_temporal.param, "cpu").to("cpu")
            self.means, self.stds = train.load_model(
                self.generator, os.path.join(MODEL_DIR, "generator.pt"), self.train_data, "cpu"
            )
        self.temporal.eval()
        self.means_latent = torch.zeros(24)
        self.stds_latent = torch.ones(24)
        self.offsets = torch.tensor(self.offsets_np, dtype=torch.float32)

    def new_drag(self, extension_losses=False):
        if extension_losses:
            DragPose = load_drag_pose_with_extension_losses()
        else:
            from drag_pose import DragPose

        return DragPose(self.generator, self.temporal, self.means_latent, self.stds_latent, "cpu", "cpu")

    def load_config(self, name):
        import json

        with open(os.path.join(CONFIG_DIR, name), "r") as fh:
            return json.load(fh)
