"""Planner facade with the reference's API (mirrors neupan/neupan.py:31-413): yaml constructor,
``forward(state, points, velocities) -> (action, info)``, lidar helpers, setters and properties.
It is host glue around the hot path: InitialPath (numpy) -> PAN (CUDA, through the C ABI).

``device`` in the yaml / ctor: 'cpu' (every example yaml of the reference) keeps the *I/O* tensors on
the host as the reference does, the computation itself always runs on the CUDA device (there is no
CPU implementation of PAN here).  ``device: 'cuda:N'`` selects the GPU and keeps ``info`` tensors there.
"""
from __future__ import annotations

from math import cos, sin

import numpy as np
import torch
import yaml

from . import configuration
from .blocks.initial_path import InitialPath
from .blocks.pan import PAN
from .configuration import np_to_tensor, tensor_to_np
from .robot import robot
from .util import file_check, time_it


def _get_transform(state):
    if state.shape == (2, 1):
        return state[0:2], np.eye(2)
    th = state[2, 0]
    return state[0:2], np.array([[cos(th), -sin(th)], [sin(th), cos(th)]])


class neupan(torch.nn.Module):
    def __init__(self, receding: int = 10, step_time: float = 0.1, ref_speed: float = 4.0, device: str = "cpu", robot_kwargs: dict = None,
                 ipath_kwargs: dict = None, pan_kwargs: dict = None, adjust_kwargs: dict = None, train_kwargs: dict = None, **kwargs) -> None:
        super().__init__()
        self.T, self.dt, self.ref_speed = receding, step_time, ref_speed
        configuration.device = torch.device(device)
        configuration.time_print = kwargs.get("time_print", False)
        self.collision_threshold = kwargs.get("collision_threshold", 0.1)
        self.cur_vel_array = np.zeros((2, self.T))
        self.robot = robot(receding, step_time, **(robot_kwargs or {}))
        self.ipath = InitialPath(receding, step_time, ref_speed, self.robot, **(ipath_kwargs or {}))
        pan_kwargs = dict(pan_kwargs or {})
        pan_kwargs["adjust_kwargs"] = adjust_kwargs
        pan_kwargs["train_kwargs"] = train_kwargs
        if configuration.device.type == "cuda":
            pan_kwargs.setdefault("device", configuration.device)
        self.dune_train_kwargs = train_kwargs
        self.pan = PAN(receding, step_time, self.robot, **pan_kwargs)
        self.device = self.pan.device
        self.info = {"stop": False, "arrive": False, "collision": False}

    @classmethod
    def init_from_yaml(cls, yaml_file, **kwargs):
        with open(file_check(yaml_file), "r") as f:
            config = yaml.safe_load(f)
        config.update(kwargs)
        for key in ("robot", "ipath", "pan", "adjust", "train"):
            config[key + "_kwargs"] = config.pop(key, dict())
        return cls(**config)

    @time_it("neupan forward")
    def forward(self, state, points, velocities=None):
        """state (3,1); points (2,N) global frame or None; velocities (2,N) or None  ->  (action (2,1), info)."""
        assert state.shape[0] >= 3
        if self.ipath.check_arrive(state):
            self.info["arrive"] = True
            return np.zeros((2, 1)), self.info
        nom_input_np = self.ipath.generate_nom_ref_state(state, self.cur_vel_array, self.ref_speed)
        nom_input_tensor = [np_to_tensor(np.asarray(n, dtype=np.float64)) for n in nom_input_np]
        pts = np_to_tensor(points) if points is not None else None
        vel = np_to_tensor(velocities) if velocities is not None else None
        opt_state_tensor, opt_vel_tensor, opt_distance_tensor = self.pan(*nom_input_tensor, pts, vel)
        opt_state_np, opt_vel_np = tensor_to_np(opt_state_tensor), tensor_to_np(opt_vel_tensor)
        self.cur_vel_array = opt_vel_np
        self.info.update(state_tensor=opt_state_tensor, vel_tensor=opt_vel_tensor, distance_tensor=opt_distance_tensor,
                         ref_state_tensor=nom_input_tensor[2], ref_speed_tensor=nom_input_tensor[3],
                         ref_state_list=[s[:, np.newaxis] for s in np.asarray(nom_input_np[2]).T],
                         opt_state_list=[s[:, np.newaxis] for s in opt_state_np.T])
        if self.check_stop():
            self.info["stop"] = True
            return np.zeros((2, 1)), self.info
        self.info["stop"] = False
        action = opt_vel_np[:, 0:1]
        if self.robot.kinematics == "omni":
            v, th = float(action[0, 0]), float(action[1, 0])
            action = np.array([[v * cos(th)], [v * sin(th)]])
            self.info["omni_linear_speed"], self.info["omni_orientation"] = v, th
        return action, self.info

    def check_stop(self):
        return float(self.min_distance) < self.collision_threshold

    # ---- lidar -> points (neupan.py:173-281; SURVEY 8f "next" #2) --------------------------------------
    def _scan_mask(self, scan, angle_range, inclusive_min):
        ranges = np.asarray(scan["ranges"], dtype=float)
        angles = np.linspace(scan["angle_min"], scan["angle_max"], len(ranges))
        lo = ranges >= scan["range_min"] if inclusive_min else ranges > scan["range_min"]
        keep = (ranges < scan["range_max"] - 0.02) & lo & (angles > angle_range[0]) & (angles < angle_range[1])
        return ranges, angles, keep

    def scan_to_point(self, state, scan, scan_offset=[0, 0, 0], angle_range=[-np.pi, np.pi], down_sample=1):
        ranges, angles, keep = self._scan_mask(scan, angle_range, inclusive_min=False)
        if not keep.any():
            return None
        local = np.vstack([ranges[keep] * np.cos(angles[keep]), ranges[keep] * np.sin(angles[keep])])
        s_trans, s_R = _get_transform(np.c_[scan_offset])
        trans, R = _get_transform(state)
        return (R @ (s_R @ local + s_trans) + trans)[:, ::down_sample]  # offset applied forward (neupan.py:216-217)

    def scan_to_point_velocity(self, state, scan, scan_offset=[0, 0, 0], angle_range=[-np.pi, np.pi], down_sample=1):
        ranges, angles, keep = self._scan_mask(scan, angle_range, inclusive_min=True)
        if not keep.any():
            return None, None
        local = np.vstack([ranges[keep] * np.cos(angles[keep]), ranges[keep] * np.sin(angles[keep])])
        scan_velocity = np.asarray(scan.get("velocity", np.zeros((2, len(ranges)))))
        s_trans, s_R = _get_transform(np.c_[scan_offset])
        trans, R = _get_transform(state)
        points = (R @ (s_R.T @ (local - s_trans)) + trans)[:, ::down_sample]  # inverse offset (neupan.py:271-274)
        return points, scan_velocity[:, keep][:, ::down_sample]

    # ---- setters / misc -------------------------------------------------------------------------------
    def train_dune(self):
        self.pan.dune_layer.train_dune(self.dune_train_kwargs)

    def reset(self):
        self.ipath.point_index = 0
        self.ipath.curve_index = 0
        self.ipath.arrive_flag = False
        self.info["stop"] = False
        self.info["arrive"] = False
        self.cur_vel_array = np.zeros_like(self.cur_vel_array)

    def set_initial_path(self, path):
        self.ipath.set_initial_path(path)

    def set_initial_path_from_state(self, state):
        self.ipath.init_check(state)

    def set_reference_speed(self, speed: float):
        self.ipath.ref_speed = speed
        self.ref_speed = speed

    def update_initial_path_from_goal(self, start, goal):
        self.ipath.update_initial_path_from_goal(start, goal)

    def update_initial_path_from_waypoints(self, waypoints):
        self.ipath.set_ipath_with_waypoints(waypoints)

    def update_adjust_parameters(self, **kwargs):
        self.pan.nrmp_layer.update_adjust_parameters_value(**kwargs)

    @property
    def min_distance(self):
        return self.pan.min_distance

    @property
    def dune_points(self):
        return self.pan.dune_points

    @property
    def nrmp_points(self):
        return self.pan.nrmp_points

    @property
    def initial_path(self):
        return self.ipath.initial_path

    @property
    def adjust_parameters(self):
        return self.pan.nrmp_layer.adjust_parameters

    @property
    def waypoints(self):
        return self.ipath.waypoints

    @property
    def opt_trajectory(self):
        return self.info["opt_state_list"]

    @property
    def ref_trajectory(self):
        return self.info["ref_state_list"]
