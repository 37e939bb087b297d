"""Drop-in for the reference's `DragPose` (python/src/drag_pose.py:12-414), one clip,
backed by the CUDA engine.  Same constructor / `set_initial_pose` / `run` signatures,
argument meaning and return shapes, so `eval_drag.py`-style drivers and `RunDrag` work
unchanged; the per-frame optimisation itself is one launch of the persistent frame
kernel (plus the temporal predictor when the reference would run it).

Differences that are deliberate and visible:
  * there is no CPU path: constructing a DragPose without a CUDA device raises;
  * `verbose` prints the loss line of drag_pose.py:361-364 but no per-phase host timers
    (the phases are fused in one kernel; use `engine.set_profiling`).
"""
from __future__ import annotations

import numpy as np
import torch

from . import model as dpm
from .engine import BatchedDragPose


def _as_np(x, dtype=np.float32):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=dtype)


def _pose_model_from(generator_model):
    """Accepts a dragposer_b200.model.PoseModel or a reference Generator_Model
    (anything with .state_dict(), .parents and .data.{mean,std}_{dqs,displacement})."""
    if isinstance(generator_model, dpm.PoseModel):
        return generator_model
    data = generator_model.data
    means = {"dqs": data.mean_dqs, "displacement": data.mean_displacement}
    stds = {"dqs": data.std_dqs, "displacement": data.std_displacement}
    return dpm.fold_generator_state(generator_model.state_dict(), means, stds, generator_model.parents)


def _temporal_from(temporal_model, means_latent, stds_latent):
    if isinstance(temporal_model, dpm.TemporalModel):
        if means_latent is not None:
            temporal_model = dpm.TemporalModel(temporal_model.sd, _as_np(means_latent).reshape(-1), _as_np(stds_latent).reshape(-1))
        return temporal_model
    return dpm.temporal_from_state(temporal_model.state_dict(), means_latent, stds_latent)


class DragPose:
    def __init__(self, generator_model, temporal_model, means_latent=None, stds_latent=None, device=None, device_gpu=None,
                 offsets=None):
        self.channels_per_joint = 4
        self.pose_model = _pose_model_from(generator_model)
        self.temporal_model = _temporal_from(temporal_model, means_latent, stds_latent)
        self.parents = list(self.pose_model.parents)
        self.device = "cpu"  # where the tensors handed back to the caller live (like the reference)
        index = 0
        for d in (device_gpu, device):
            if d is not None and str(d).startswith("cuda") and ":" in str(d):
                index = int(str(d).split(":")[1])
        self._cuda_index = index
        self._engine = None
        self._offsets = None if offsets is None else _as_np(offsets).reshape(22, 3)
        self.means_dqs = torch.from_numpy(self.pose_model.mean_q.copy()).unsqueeze(-1)
        self.stds_dqs = torch.from_numpy(self.pose_model.std_q.copy()).unsqueeze(-1)
        self.means_latent = torch.from_numpy(self.temporal_model.means_latent.copy())
        self.stds_latent = torch.from_numpy(self.temporal_model.stds_latent.copy())
        self._pending = None
        self.last_iterations = None
        self.last_losses = None

    # -- engine is created lazily because the reference only learns the offsets in run()
    def _ensure_engine(self, offsets):
        offsets = _as_np(offsets).reshape(22, 3)
        if self._engine is None or not np.array_equal(offsets, self._offsets):
            if self._engine is not None:
                saved = self._engine.state()
                self._engine.close()
            else:
                saved = None
            self._offsets = offsets
            self._engine = BatchedDragPose(self.pose_model, offsets, self.temporal_model, 1, device=self._cuda_index)
            if saved is not None:
                self._engine.set_initial_state(saved["latent"], saved["global_pos"], saved["global_rot"], np.zeros((1, 6)))
                self._engine.set_ring_buffers(saved["latent_buf"], saved["disp_buf"], saved["height_buf"])
        if self._pending is not None:
            self._engine.set_initial_state(*self._pending)
            self._pending = None
        return self._engine

    def set_initial_pose(self, initial_pose, init_global_pos, initial_global_rot, initial_heights, eps=None):
        """initial_pose (1,176,W) standardised dual quats (the last frame is encoded),
        init_global_pos (1,3,1), initial_global_rot (1,4,1), initial_heights (6,).
        The latent is mu + eps * exp(0.5 logvar) with eps ~ torch.randn (autoencoder.py:19-27)."""
        dqs = _as_np(initial_pose).reshape(1, 176, -1)[..., -1]
        mu, logvar = self.pose_model.encode_np(dqs)
        if eps is None:
            eps = torch.randn(1, dpm.LATENT).numpy()
        latent = mu + _as_np(eps).reshape(1, -1) * np.exp(np.float32(0.5) * logvar)
        self.set_initial_latent(latent, init_global_pos, initial_global_rot, initial_heights)

    def set_initial_latent(self, latent, init_global_pos, initial_global_rot, initial_heights):
        self._pending = (_as_np(latent).reshape(1, 24), _as_np(init_global_pos).reshape(1, 3),
                         _as_np(initial_global_rot).reshape(1, 4), _as_np(initial_heights).reshape(1, 6))
        if self._engine is not None:
            self._engine.set_initial_state(*self._pending)
            self._pending = None
        self.current_index = 0

    # -- carried state, read back on demand
    def _state(self):
        if self._pending is not None:
            lat, gp, gr, ht = self._pending
            return dict(latent=lat, global_pos=gp, global_rot=gr, latent_buf=np.tile(lat[:, None], (1, 60, 1)),
                        disp_buf=np.zeros((1, 60, 3), np.float32), height_buf=np.tile(ht[:, None], (1, 60, 1)))
        return self._engine.state()

    @property
    def current_global_pos(self):
        return torch.from_numpy(self._state()["global_pos"].reshape(1, 3, 1).copy())

    @current_global_pos.setter
    def current_global_pos(self, value):
        gp = _as_np(value).reshape(1, 3)
        if self._pending is not None:
            self._pending = (self._pending[0], gp, self._pending[2], self._pending[3])
        else:
            self._engine.set_global_pos(gp)

    @property
    def current_global_rot(self):
        return torch.from_numpy(self._state()["global_rot"].reshape(1, 4).copy())

    @property
    def latent(self):
        return torch.from_numpy(self._state()["latent"].reshape(1, 24).copy())

    @property
    def latent_buffer(self):
        return torch.from_numpy(self._state()["latent_buf"][0].copy())

    @property
    def displacement_buffer(self):
        return torch.from_numpy(self._state()["disp_buf"][0].copy())

    @property
    def heights_buffer(self):
        return torch.from_numpy(self._state()["height_buf"][0].copy())

    def run(self, target_ee_pos, target_ee_rot, mask_joints, weights_joints, offsets, stop_eps_pos=1e-2, stop_eps_rot=1e-2,
            max_iter=100, min_loss_incr=0.00001, learning_rate=1e-3, lambda_rot=1, lambda_temporal=1,
            temporal_future_window=60, height_indices=(0, 4, 8, 13, 17, 21), joint_adjustment_indices=None,
            joint_adjustment_weight=0.01, verbose=False, extension_losses=0, floor_level=0.0):
        """extension_losses / floor_level (not in the reference signature): bit mask re-enabling the reference's commented-out
        "Additional Losses" (drag_pose.py:129-183), see engine.EXT_*; 0 = the shipped behaviour."""
        if list(height_indices) != [0, 4, 8, 13, 17, 21]:
            raise ValueError("height_indices other than train_temporal.param['height_indices'] are not supported")
        assert temporal_future_window % dpm.SAMPLE_STEP == 0  # drag_pose.py:236
        eng = self._ensure_engine(offsets)
        joints = _as_np(mask_joints, np.int32).reshape(-1)
        E = joints.shape[0]
        pose, gpos = eng.run(_as_np(target_ee_pos).reshape(1, E, 3), _as_np(target_ee_rot).reshape(1, E, 3, 3), joints,
                             _as_np(weights_joints).reshape(E, 2), stop_eps_pos=stop_eps_pos, stop_eps_rot=stop_eps_rot,
                             max_iter=max_iter, min_loss_incr=min_loss_incr, learning_rate=learning_rate, lambda_rot=lambda_rot,
                             lambda_temporal=lambda_temporal, temporal_future_window=temporal_future_window,
                             joint_adjustment_indices=joint_adjustment_indices, joint_adjustment_weight=joint_adjustment_weight,
                             extension_losses=extension_losses, floor_level=floor_level)
        iters, losses = eng.frame_stats()
        self.last_iterations, self.last_losses = int(iters[0]), losses[0]
        if verbose:
            print(f"Loss sqrt(Pos): {np.sqrt(losses[0, 0]):.5f} // Loss Rot: {losses[0, 1]:.5f} // "
                  f"Loss Temporal: {losses[0, 2]:.5f} // Iter: {int(iters[0])}")
        return torch.from_numpy(pose[0]), torch.from_numpy(gpos[0])

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None
