"""Drop-in for the reference's `RunDrag` session object (python/src/run_drag.py:14-176):
the stateful wrapper that DragPoserDLL drives through the C ABI.  Same method names,
argument meaning and in-place result convention (numpy arrays written with copyto)."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import model as dpm
from . import rotations as rot
from .bvh import Bvh
from .drag_pose import DragPose


class RunDrag:
    def __init__(self, seed=2222):
        torch.manual_seed(seed)  # run_drag.py:17-19 (train.param["seed"])
        np.random.seed(seed)
        self.device = "cpu"
        self.drag = None

    def set_reference_skeleton(self, bvh_path):
        self.parents, offsets = Bvh(bvh_path).skeleton()
        self.offsets = torch.from_numpy(offsets)
        return len(self.parents)

    def load_models(self, model_path, temporal=None):
        """Directory with generator.pt + data.pt + temporal.pt (run_drag.py:40-59), or a folded .npz with temporal.pt next
        to it.  Like the reference, a missing temporal.pt raises; `temporal` = an explicit TemporalModel (tests, benches:
        the seeded random-init predictor) overrides the file."""
        self.pose_model = dpm.load_pose_model(model_path, self.parents)
        tdir = model_path if os.path.isdir(model_path) else os.path.dirname(model_path)
        self.temporal_model = temporal if temporal is not None else dpm.load_temporal_model(tdir)
        self.means = {"dqs": self.pose_model.mean_dqs}
        self.stds = {"dqs": self.pose_model.std_dqs}

    def set_mask_and_weights(self, mask, weights):
        mask, weights = np.asarray(mask), np.asarray(weights)
        assert len(mask) == len(self.parents)
        assert len(weights) == len(self.parents)
        assert weights.shape[1] == 2  # position and rotation
        self.mask_indices = np.nonzero(mask)[0].astype(np.int32)
        if self.mask_indices.size < 2:
            # the reference's nonzero().squeeze() turns a single tracker into a 0-d index and fails
            raise ValueError("at least two trackers are required")
        self.weights = weights[self.mask_indices].astype(np.float32)
        return len(self.mask_indices)

    def init_drag_pose(self, initial_global_pos, initial_global_rot, eps=None, initial_latent=None):
        """Encodes the zero (= mean) standardised pose, heights 0 (run_drag.py:77-96).  `eps` / `initial_latent` reproduce a
        recorded reparameterisation draw (the reference takes it from torch's RNG stream)."""
        if self.drag is not None:
            self.drag.close()
        self.drag = DragPose(self.pose_model, self.temporal_model, offsets=self.offsets)
        gp, gr = np.asarray(initial_global_pos).reshape(1, 3, 1), np.asarray(initial_global_rot).reshape(1, 4, 1)
        if initial_latent is not None:
            self.drag.set_initial_latent(initial_latent, gp, gr, np.zeros(6, np.float32))
        else:
            self.drag.set_initial_pose(np.zeros((1, len(self.parents) * 8, 1), np.float32), gp, gr, np.zeros(6, np.float32), eps=eps)

    def set_optim_params(self, stop_eps_pos, stop_eps_rot, max_iter, lr):
        self.stop_eps_pos, self.stop_eps_rot, self.max_iter, self.learning_rate = stop_eps_pos, stop_eps_rot, max_iter, lr

    def set_lambdas(self, lambda_rot, lambda_temporal, temporal_future_window):
        self.lambda_rot, self.lambda_temporal, self.temporal_future_window = lambda_rot, lambda_temporal, temporal_future_window

    def set_global_pos(self, global_pos):
        self.drag.current_global_pos = np.asarray(global_pos, np.float32).reshape(1, 3)

    def drag_pose(self, target_ee_pos, target_ee_rot, result_pose, result_global_pos):
        """target_ee_pos (E,3), target_ee_rot (E,4) wxyz quaternions; writes local quaternions
        (J,4) into result_pose and the root position into result_global_pos (1,3)."""
        tgt_rot = rot.to_matrix(np.asarray(target_ee_rot, np.float32))
        res_pose, res_gpos = self.drag.run(
            target_ee_pos=np.asarray(target_ee_pos, np.float32), target_ee_rot=tgt_rot, mask_joints=self.mask_indices,
            weights_joints=self.weights, offsets=self.offsets, stop_eps_pos=self.stop_eps_pos, stop_eps_rot=self.stop_eps_rot,
            max_iter=self.max_iter, learning_rate=self.learning_rate, lambda_rot=self.lambda_rot,
            lambda_temporal=self.lambda_temporal, temporal_future_window=self.temporal_future_window,
            joint_adjustment_indices=None, verbose=False)
        q = res_pose.numpy() * self.pose_model.std_q + self.pose_model.mean_q
        rots = rot.from_root_quat(q.reshape(1, -1, 4), self.parents)
        np.copyto(result_pose, rots.reshape(-1, 4))
        result_global_pos[0, :] = res_gpos.numpy()
