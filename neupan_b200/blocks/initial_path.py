"""InitialPath: reference trajectory bookkeeping in front of the hot path (SURVEY.md 8f "next" #1).

Mirrors the interface of neupan/blocks/initial_path.py:28-498 (same constructor keywords, attributes
and methods, same nominal / reference construction in ``generate_nom_ref_state``), host-side numpy.
The reference delegates curve generation to the third-party ``gctl.curve_generator`` (pinned
gctl==1.2, absent here); this module uses gctl when it is importable and otherwise its own generator:
``line`` (points every ``interval`` along each segment), ``dubins`` (shortest of the six Dubins
words at ``min_radius``) and ``reeds`` (shortest Reeds-Shepp word, ``reeds_shepp.py``; backward
segments carry gear -1).  Point placement of the built-in generator is not
verified against gctl (it could not be run here) -- only the spacing / heading conventions the
reference relies on are kept: points are (4,1) columns [x, y, theta, gear].
"""
from __future__ import annotations

import math
from math import cos, inf, sin, sqrt, tan

import numpy as np

from ..util import WrapToPi


def _distance(a, b) -> float:
    return sqrt((a[0, 0] - b[0, 0]) ** 2 + (a[1, 0] - b[1, 0]) ** 2)


class CurveGenerator:
    """Stand-in for gctl.curve_generator().generate_curve(style, waypoints, interval, min_radius, include_gear)."""

    def generate_curve(self, curve_style, way_points, step_size=0.1, min_radius=1.0, include_gear=False, **kwargs):
        pts = []
        for a, b in zip(way_points[:-1], way_points[1:]):
            a, b = np.asarray(a, float).reshape(-1), np.asarray(b, float).reshape(-1)
            if curve_style == "line":
                seg = self._line(a, b, step_size)
            elif curve_style == "dubins":
                seg = self._dubins(a, b, step_size, max(min_radius, 1e-6))
            elif curve_style == "reeds":
                from .reeds_shepp import sample_path

                seg = sample_path(a, b, step_size, max(min_radius, 1e-6))
            else:
                raise ValueError(f"curve_style '{curve_style}' is not one of line, dubins, reeds")
            if pts and seg:
                seg = seg[1:]  # the joint waypoint is already there
            pts += seg
        out = []
        for pt in pts:
            x, y, th = pt[0], pt[1], pt[2]
            gear = pt[3] if len(pt) > 3 else 1.0  # only Reeds-Shepp curves drive backwards
            col = np.array([[x], [y], [th], [gear]]) if include_gear else np.array([[x], [y], [th]])
            out.append(col)
        return out

    @staticmethod
    def _line(a, b, step):
        d = math.hypot(b[0] - a[0], b[1] - a[1])
        th = math.atan2(b[1] - a[1], b[0] - a[0]) if d > 0 else (a[2] if len(a) > 2 else 0.0)
        n = max(int(math.floor(d / step)), 0) if step > 0 else 0
        pts = [(a[0] + (b[0] - a[0]) * (k * step / d), a[1] + (b[1] - a[1]) * (k * step / d), th) for k in range(n + 1)] if d > 0 else [(a[0], a[1], th)]
        if d > 0 and d - n * step > 1e-9:
            pts.append((b[0], b[1], th))
        return pts

    @staticmethod
    def _dubins(a, b, step, r):
        """Shortest Dubins path from pose a to pose b with turning radius r, sampled every `step`."""
        dx, dy = (b[0] - a[0]) / r, (b[1] - a[1]) / r
        D = math.hypot(dx, dy)
        th = math.atan2(dy, dx)
        al, be = (a[2] - th) % (2 * math.pi), (b[2] - th) % (2 * math.pi)
        sa, sb, ca, cb, cab = sin(al), sin(be), cos(al), cos(be), cos(al - be)
        mod = lambda x: x % (2 * math.pi)
        cands = []
        p2 = 2 + D * D - 2 * cab + 2 * D * (sa - sb)
        if p2 >= 0:  # LSL
            t = math.atan2(cb - ca, D + sa - sb)
            cands.append(("LSL", mod(-al + t), sqrt(p2), mod(be - t)))
        p2 = 2 + D * D - 2 * cab + 2 * D * (sb - sa)
        if p2 >= 0:  # RSR
            t = math.atan2(ca - cb, D - sa + sb)
            cands.append(("RSR", mod(al - t), sqrt(p2), mod(-be + t)))
        p2 = -2 + D * D + 2 * cab + 2 * D * (sa + sb)
        if p2 >= 0:  # LSR
            p = sqrt(p2)
            t = math.atan2(-ca - cb, D + sa + sb) - math.atan2(-2.0, p)
            cands.append(("LSR", mod(-al + t), p, mod(-mod(be) + t)))
        p2 = D * D - 2 + 2 * cab - 2 * D * (sa + sb)
        if p2 >= 0:  # RSL
            p = sqrt(p2)
            t = math.atan2(ca + cb, D - sa - sb) - math.atan2(2.0, p)
            cands.append(("RSL", mod(al - t), p, mod(be - t)))
        q = (6 - D * D + 2 * cab + 2 * D * (sa - sb)) / 8
        if abs(q) <= 1:  # RLR
            p = mod(2 * math.pi - math.acos(q))
            t = mod(al - math.atan2(ca - cb, D - sa + sb) + p / 2)
            cands.append(("RLR", t, p, mod(al - be - t + p)))
        q = (6 - D * D + 2 * cab + 2 * D * (sb - sa)) / 8
        if abs(q) <= 1:  # LRL
            p = mod(2 * math.pi - math.acos(q))
            t = mod(-al - math.atan2(ca - cb, D + sa - sb) + p / 2)
            cands.append(("LRL", t, p, mod(mod(be) - al - t + p)))
        word, t, p, q_ = min(cands, key=lambda c: c[1] + c[2] + c[3])
        lengths = [t * r, p * r, q_ * r]
        total = sum(lengths)
        pts = []
        n = max(int(math.floor(total / step)), 0)
        for k in list(range(n + 1)) + ([None] if total - n * step > 1e-9 else []):
            s = total if k is None else k * step
            x, y, h = a[0], a[1], a[2]
            rem = s
            for seg_type, L in zip(word, lengths):
                l = min(rem, L)
                if seg_type == "S":
                    x += l * cos(h); y += l * sin(h)
                else:
                    sgn = 1.0 if seg_type == "L" else -1.0
                    x += sgn * r * (sin(h + sgn * l / r) - sin(h)); y += -sgn * r * (cos(h + sgn * l / r) - cos(h))
                    h += sgn * l / r
                rem -= l
                if rem <= 0:
                    break
            pts.append((x, y, WrapToPi(h)))
        return pts


def _make_generator():
    try:
        import gctl  # the reference's dependency, when present

        if getattr(gctl, "__file__", None) is None:  # an in-memory placeholder (e.g. the test loader's stub), not the package
            raise ImportError("gctl is not installed")
        return gctl.curve_generator()
    except Exception:
        return CurveGenerator()


class InitialPath:
    def __init__(self, receding, step_time, ref_speed, robot, waypoints=None, loop=False, curve_style="line", **kwargs) -> None:
        self.T, self.dt, self.ref_speed, self.robot = receding, step_time, ref_speed, robot
        self.waypoints = self.trans_to_np_list(waypoints)
        self.loop, self.curve_style = loop, curve_style
        self.min_radius = kwargs.get("min_radius", self.default_turn_radius())
        self.interval = kwargs.get("interval", self.dt * self.ref_speed)
        self.arrive_threshold = kwargs.get("arrive_threshold", 0.1)
        self.close_threshold = kwargs.get("close_threshold", 0.1)
        self.ind_range = kwargs.get("ind_range", 10)
        self.arrive_index_threshold = kwargs.get("arrive_index_threshold", 1)
        self.arrive_flag = False
        self.cg = _make_generator()
        self.initial_path = None
        self.curve_list, self.curve_index, self.point_index = [], 0, 0

    # ---- nominal / reference trajectories (initial_path.py:68-126) --------------------------------
    def generate_nom_ref_state(self, state: np.ndarray, cur_vel_array: np.ndarray, ref_speed: float):
        """nom_s: rollout of the previous optimal controls from `state`; ref_s: points `ref_speed*dt`
        further along the current curve per step (heading unwrapped towards nom_s); ref_us = gear*ref_speed."""
        state = state[:3]
        ref_state = self.cur_point[0:3].copy()
        ref_index = self.point_index
        pre_state = state.copy()
        state_pre_list, state_ref_list = [pre_state], [ref_state]
        assert self.cur_point.shape[0] >= 4
        gear_list = [self.cur_point[-1, 0]] * self.T
        ref_speed_forward = ref_speed * self.dt
        last = len(self.cur_curve) - 1
        for t in range(self.T):
            pre_state = self.motion_predict_model(pre_state, cur_vel_array[:, t:t + 1], self.robot.L, self.dt)
            state_pre_list.append(pre_state)
            if ref_speed_forward >= self.interval:
                ref_index += int(ref_speed_forward / self.interval)
                if ref_index > last:
                    ref_index = last
                    gear_list[t] = 0
                ref_state = self.cur_curve[ref_index][0:3]
            else:
                ref_state, ref_index = self.find_interaction_point(ref_state, ref_index, ref_speed_forward)
                if ref_index > last:
                    gear_list[t] = 0
            # NB: like the reference this writes through to the stored path point (a view), initial_path.py:111-112
            ref_state[2, 0] = pre_state[2, 0] + WrapToPi(ref_state[2, 0] - pre_state[2, 0])
            state_ref_list.append(ref_state)
        nom_s = np.hstack(state_pre_list)
        ref_s = np.hstack(state_ref_list)
        ref_us = np.array(gear_list) * ref_speed
        return nom_s, cur_vel_array, ref_s, ref_us

    def motion_predict_model(self, robot_state, vel, wheel_base, sample_time):
        phi, v, w = robot_state[2, 0], vel[0, 0], vel[1, 0]
        kin = self.robot.kinematics
        if kin == "acker":
            ds = np.array([[v * cos(phi)], [v * sin(phi)], [v * tan(w) / wheel_base]])
        elif kin == "diff":
            ds = np.array([[v * cos(phi)], [v * sin(phi)], [w]])
        else:  # omni: (speed, heading)
            ds = np.array([[v * cos(w)], [v * sin(w)], [0.0]])
        return robot_state[0:3] + ds * sample_time

    # ---- path bookkeeping --------------------------------------------------------------------------
    def set_initial_path(self, path):
        self.initial_path = path
        self.interval = self.cal_average_interval(path)
        self.split_path_with_gear()
        self.curve_index = self.point_index = 0

    def cal_average_interval(self, path):
        if len(path) < 2:
            return 0
        dist_sum = 0.0  # sequential additions like initial_path.py:157-162 (the builtin sum() is compensated and can differ in the
        for p, q in zip(path, path[1:]):  # last bit, which flips `ref_speed * dt >= interval` when the spacing equals ref_speed * dt)
            dist_sum += math.hypot(float(np.asarray(q[0] - p[0]).reshape(-1)[0]), float(np.asarray(q[1] - p[1]).reshape(-1)[0]))
        return dist_sum / (len(path) - 1)

    def closest_point(self, state, threshold=0.1, ind_range=10):
        min_dis = inf
        start, end = max(self.point_index, 0), min(self.point_index + ind_range, len(self.cur_curve))
        for index in range(start, end):
            dis = _distance(state[0:2], self.cur_curve[index][0:2])
            if dis < min_dis:
                min_dis = dis
                self.point_index = index
                if dis < threshold:
                    break
        return min_dis

    def find_interaction_point(self, ref_state, ref_index, length):
        circle = np.squeeze(ref_state[0:2])
        while True:
            if ref_index > len(self.cur_curve) - 2:
                end_point = self.cur_curve[-1]
                end_point[2] = WrapToPi(end_point[2, 0])
                return end_point[0:3], ref_index
            cur_point, next_point = self.cur_curve[ref_index], self.cur_curve[ref_index + 1]
            hit = self.range_cir_seg(circle, length, [np.squeeze(cur_point[0:2]), np.squeeze(next_point[0:2])])
            if hit is not None:
                diff = WrapToPi(next_point[2, 0] - cur_point[2, 0])
                theta = WrapToPi(cur_point[2, 0] + diff / 2)
                return np.append(hit, theta).reshape((3, 1)), ref_index
            ref_index += 1

    def range_cir_seg(self, circle, r, segment):
        sp, ep = segment
        d = ep - sp
        if np.linalg.norm(d) == 0:
            return None
        f = sp - circle
        a, b, c = d @ d, 2 * f @ d, f @ f - r ** 2
        disc = b ** 2 - 4 * a * c
        if disc < 0:
            return None
        t2 = (-b + sqrt(disc)) / (2 * a)
        return sp + t2 * d if 0 <= t2 <= 1 else None

    def check_arrive(self, state):
        self.init_check(state)
        self.closest_point(state, self.close_threshold, self.ind_range)
        if self.check_curve_arrive(state, self.arrive_threshold, self.arrive_index_threshold):
            if self.curve_index + 1 >= self.curve_number:
                if self.loop:
                    self.curve_index = self.point_index = 0
                    print("Info: loop, reset the path")
                    return False
                if not self.arrive_flag:
                    print("Info: arrive at the end of the path")
                    self.arrive_flag = True
                return True
            self.curve_index += 1
            self.point_index = 0
        return False

    def check_curve_arrive(self, state, arrive_threshold=0.1, arrive_index_threshold=2):
        arrive_distance = np.linalg.norm(state[0:2] - self.cur_curve[-1][0:2])
        return arrive_distance < arrive_threshold and self.point_index >= (len(self.cur_curve) - arrive_index_threshold - 2)

    def split_path_with_gear(self):
        self.curve_list, current, gear = [], [], self.initial_path[0][-1]
        for point in self.initial_path:
            if point[-1] != gear:
                self.curve_list.append(current)
                current, gear = [], point[-1]
            current.append(point)
        if current:
            self.curve_list.append(current)

    def _regenerate(self, waypoints):
        self.initial_path = self.cg.generate_curve(self.curve_style, waypoints, self.interval, self.min_radius, True)
        if self.curve_style == "line":
            self._ensure_consistent_angles()
        self.split_path_with_gear()
        self.curve_index = self.point_index = 0

    def init_path_with_state(self, state):
        assert len(self.waypoints) > 0, "Error: waypoints are not set"
        if isinstance(self.waypoints, list):
            self.waypoints = [state] + self.waypoints
        elif isinstance(self.waypoints, np.ndarray):
            self.waypoints = np.vstack([state, self.waypoints])
        if self.loop:
            self.waypoints = self.waypoints + [self.waypoints[0]]
        self.initial_path = self.cg.generate_curve(self.curve_style, self.waypoints, self.interval, self.min_radius, True)
        if self.curve_style == "line":
            self._ensure_consistent_angles()

    def init_check(self, state):
        if self.initial_path is None:
            print("initial path is not set, generate path with the current state")
            self.set_ipath_with_state(state)

    def set_ipath_with_state(self, state):
        self.init_path_with_state(state[0:3])
        self.split_path_with_gear()
        self.curve_index = self.point_index = 0

    def update_initial_path_from_goal(self, start, goal):
        waypoints = [start, goal, start] if self.loop else [start, goal]
        self._regenerate(waypoints)
        self.waypoints = waypoints

    def set_ipath_with_waypoints(self, waypoints):
        self._regenerate(waypoints)
        self.waypoints = waypoints

    @property
    def cur_waypoints(self):
        return self.waypoints

    @property
    def cur_curve(self):
        return self.curve_list[self.curve_index]

    @property
    def cur_point(self):
        return self.cur_curve[self.point_index]

    @property
    def curve_number(self):
        return len(self.curve_list)

    def default_turn_radius(self):
        if self.robot.kinematics == "acker":
            return self.robot.L / tan(float(np.asarray(self.robot.max_speed).reshape(-1)[1]))
        return 0.0

    def _ensure_consistent_angles(self):
        if self.initial_path is None or len(self.initial_path) < 2:
            return
        for cur, nxt in zip(self.initial_path[:-1], self.initial_path[1:]):
            cur[2, 0] = math.atan2(nxt[1, 0] - cur[1, 0], nxt[0, 0] - cur[0, 0])
        self.initial_path[-1][2, 0] = self.initial_path[-2][2, 0]

    def trans_to_np_list(self, point_list):
        if point_list is None:
            return []
        return [np.c_[p] if isinstance(p, list) else p for p in point_list]
