"""DUNE layer facade (mirrors neupan/blocks/dune.py:29-216).

Holds what the reference's DUNE holds -- the ObsPointNet parameters, G, h, ``min_distance``,
``points`` -- and loads the same checkpoints.  Its per-step computation (dune.py:58-127) happens
inside the CUDA DUNE kernel driven by PAN; this class owns no compute.
"""
from __future__ import annotations

from math import inf
from typing import Optional

import os

import numpy as np
import torch

from ..util import file_check
from .obs_point_net import ObsPointNet


class DUNE(torch.nn.Module):
    def __init__(self, receding: int = 10, checkpoint=None, robot=None, dune_max_num: int = 100, train_kwargs: Optional[dict] = None) -> None:
        super().__init__()
        if robot is None:
            raise ValueError("robot parameter is required and cannot be None")  # dune.py:37-38
        self.T = receding
        self.max_num = dune_max_num
        self.robot = robot
        self.G = torch.from_numpy(np.asarray(robot.G)).float()
        self.h = torch.from_numpy(np.asarray(robot.h)).float()
        self.edge_dim = self.G.shape[0]
        self.state_dim = self.G.shape[1]
        self.model = ObsPointNet(2, self.edge_dim)
        self.load_model(checkpoint, train_kwargs)
        self.obstacle_points = None
        self.min_distance = inf

    def load_model(self, checkpoint: Optional[str] = None, train_kwargs: Optional[dict] = None):
        """dune.py:131-170 without the interactive prompts: a missing checkpoint raises
        FileNotFoundError unless ``train_kwargs['direct_train']`` is set (weights stay at init)."""
        if checkpoint is None or str(checkpoint) == "None":
            if train_kwargs and train_kwargs.get("direct_train", False):
                return
            raise FileNotFoundError("DUNE checkpoint is required (pan.dune_checkpoint); training on demand is not part of the hot path")
        try:
            path = file_check(checkpoint)
        except FileNotFoundError:
            # dune.py:146-170: a named but missing checkpoint is not fatal when the caller asked to train directly
            if train_kwargs and train_kwargs.get("direct_train", False):
                print(f"checkpoint {checkpoint} not found; direct_train is set: weights stay at their initial values until train_dune()")
                return
            raise
        if str(checkpoint).endswith(".npz"):
            z = np.load(path)
            sd = {k: torch.from_numpy(z[k]) for k in z.files if k.startswith("MLP.")}
        else:
            sd = torch.load(path, map_location=torch.device("cpu"))
        self.abs_checkpoint_path = path
        self.model.load_state_dict(sd)
        self.model.eval()

    def train_dune(self, train_kwargs):
        """dune.py:173-182: trains ``self.model`` and saves ``model/<model_name>/model_<epoch>.pth`` next to the running script.
        The native PAN handle keeps the weights it was created with: construct the planner from the new checkpoint to use it."""
        import sys

        from .dune_train import DUNETrain

        train_kwargs = dict(train_kwargs or {})
        model_name = train_kwargs.pop("model_name", getattr(self.robot, "name", "robot"))
        train_kwargs.pop("direct_train", None)
        base = train_kwargs.pop("checkpoint_dir", os.path.join(sys.path[0], "model"))
        path, k = os.path.join(base, model_name), 0
        while os.path.exists(path):  # util.repeat_mk_dirs: append a counter instead of overwriting
            k += 1
            path = os.path.join(base, f"{model_name}_{k}")
        self.train_model = DUNETrain(self.model, self.G, self.h, path)
        self.full_model_name = self.train_model.start(**train_kwargs)
        print("Complete Training. The model is saved in " + str(self.full_model_name))
        self.weights_version = getattr(self, "weights_version", 0) + 1  # PAN re-packs the native weight image on the next forward
        return self.full_model_name

    @property
    def points(self):
        return self.obstacle_points
