"""CPU restatement of the per-control-step part of ``InitialPath`` (TEST INFRASTRUCTURE ONLY, like the rest of oracle/).

Follows ``InitialPath.check_arrive`` / ``closest_point`` / ``check_curve_arrive`` (neupan/blocks/initial_path.py:166-183,
251-292), ``generate_nom_ref_state`` (:68-126), ``find_interaction_point`` / ``range_cir_seg`` (:185-249), the three
motion models (:386-444), ``cal_average_interval`` (:146-164) and ``split_path_with_gear`` (:294-317): the work
``neupan.forward`` does before PAN at every control step (neupan/neupan.py:114-121).  Curve generation from waypoints
(gctl, third party, absent here) is NOT part of it: paths enter as explicit point lists, as through
``neupan.set_initial_path`` (neupan.py:296-303).

The reference keeps the path as a list of (4,1) float64 arrays [x, y, theta, gear] and hands out numpy VIEWS of them:
``ref_state = self.cur_curve[ref_index][0:3]`` (:99) followed by ``ref_state[2, 0] = ...`` (:112) rewrites the heading of
the stored path point, entries of ``state_ref_list`` that alias the same point change together (the clamped tail of a
curve), and ``find_interaction_point`` wraps the heading of the last point in place (:191-192).  These effects persist
across control steps.  The restatement uses the same data structure and the same view semantics, so it inherits all of
them; tests/test_ipath.py pins it to the reference class itself and to golden vectors it produced.
"""
import math
from math import cos, inf, sin, sqrt, tan

import numpy as np


def wrap_to_pi(rad):
    """util.WrapToPi (neupan/util/__init__.py:98-120)."""
    while rad > math.pi:
        rad = rad - 2 * math.pi
    while rad < -math.pi:
        rad = rad + 2 * math.pi
    return rad


def _distance(p1, p2):
    """util.distance (neupan/util/__init__.py:122-133)."""
    return sqrt((p1[0, 0] - p2[0, 0]) ** 2 + (p1[1, 0] - p2[1, 0]) ** 2)


class OracleInitialPath:
    def __init__(self, T, dt, kinematics, L=None, loop=False, arrive_threshold=0.1, close_threshold=0.1, ind_range=10, arrive_index_threshold=1):
        self.T, self.dt, self.kinematics, self.L, self.loop = T, dt, kinematics, L, loop
        self.arrive_threshold, self.close_threshold = arrive_threshold, close_threshold
        self.ind_range, self.arrive_index_threshold = ind_range, arrive_index_threshold
        self.arrive_flag = False
        self.curve_list, self.curve_index, self.point_index, self.interval = [], 0, 0, 0.0

    # ---- path set-up (initial_path.py:128-164, 294-317) ---------------------------------------------
    def set_initial_path(self, path):
        self.initial_path = path
        n = len(path)
        dist_sum = 0.0
        for p1, p2 in zip(path, path[1:]):
            dist_sum += math.hypot(p2[0, 0] - p1[0, 0], p2[1, 0] - p1[1, 0])
        self.interval = dist_sum / (n - 1) if n >= 2 else 0
        self.curve_list, cur, gear = [], [], path[0][-1]
        for point in path:
            if point[-1] != gear:
                self.curve_list.append(cur)
                cur, gear = [], point[-1]
            cur.append(point)
        if cur:
            self.curve_list.append(cur)
        self.curve_index = self.point_index = 0

    @property
    def cur_curve(self):
        return self.curve_list[self.curve_index]

    # ---- check_arrive (initial_path.py:251-292) ----------------------------------------------------------
    def closest_point(self, state):
        min_dis, cur = inf, self.point_index
        for index in range(max(cur, 0), min(cur + self.ind_range, len(self.cur_curve))):
            dis = _distance(state[0:2], self.cur_curve[index][0:2])
            if dis < min_dis:
                min_dis = dis
                self.point_index = index
                if dis < self.close_threshold:
                    break
        return min_dis

    def check_arrive(self, state):
        self.closest_point(state)
        final = self.cur_curve[-1][0:2]
        arrive = (np.linalg.norm(state[0:2] - final) < self.arrive_threshold
                  and self.point_index >= (len(self.cur_curve) - self.arrive_index_threshold - 2))
        if arrive:
            if self.curve_index + 1 >= len(self.curve_list):
                if self.loop:
                    self.curve_index = self.point_index = 0
                    return False
                self.arrive_flag = True
                return True
            self.curve_index += 1
            self.point_index = 0
        return False

    # ---- generate_nom_ref_state (initial_path.py:68-126) ------------------------------------------------
    def _predict(self, s, vel):
        v, w = vel[0, 0], vel[1, 0]
        if self.kinematics == "acker":
            ds = np.array([[v * cos(s[2, 0])], [v * sin(s[2, 0])], [v * tan(w) / self.L]])
            return s + ds * self.dt
        if self.kinematics == "diff":
            ds = np.array([[v * cos(s[2, 0])], [v * sin(s[2, 0])], [w]])
            return s + ds * self.dt
        return s + self.dt * np.array([[v * cos(w)], [v * sin(w)], [0]])  # omni (:432-444)

    def _range_cir_seg(self, circle, r, sp, ep):
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

    def _find_interaction_point(self, ref_state, ref_index, length):
        circle = np.squeeze(ref_state[0:2])
        while True:
            if ref_index > len(self.cur_curve) - 2:
                end_point = self.cur_curve[-1]
                end_point[2] = wrap_to_pi(end_point[2])  # in place (:191-192)
                return end_point[0:3], ref_index
            cur, nxt = self.cur_curve[ref_index], self.cur_curve[ref_index + 1]
            pt = self._range_cir_seg(circle, length, np.squeeze(cur[0:2]), np.squeeze(nxt[0:2]))
            if pt is not None:
                diff = wrap_to_pi(nxt[2, 0] - cur[2, 0])
                theta = wrap_to_pi(cur[2, 0] + diff / 2)
                return np.append(pt, theta).reshape((3, 1)), ref_index
            ref_index += 1

    def generate_nom_ref_state(self, state, cur_vel_array, ref_speed):
        state = state[:3]
        cur_point = self.cur_curve[self.point_index]
        ref_state, ref_index, pre = cur_point[0:3].copy(), self.point_index, state.copy()
        pre_list, ref_list = [pre], [ref_state]
        gear_list = [cur_point[-1, 0]] * self.T
        fwd = ref_speed * self.dt
        for t in range(self.T):
            pre = self._predict(pre, cur_vel_array[:, t:t + 1])
            pre_list.append(pre)
            if fwd >= self.interval:
                ref_index = ref_index + int(fwd / self.interval)
                if ref_index > len(self.cur_curve) - 1:
                    ref_index = len(self.cur_curve) - 1
                    gear_list[t] = 0
                ref_state = self.cur_curve[ref_index][0:3]  # a VIEW of the stored point (:99)
            else:
                ref_state, ref_index = self._find_interaction_point(ref_state, ref_index, fwd)
                if ref_index > len(self.cur_curve) - 1:
                    gear_list[t] = 0
            diff = ref_state[2, 0] - pre[2, 0]
            ref_state[2, 0] = pre[2, 0] + wrap_to_pi(diff)  # writes through the view (:112)
            ref_list.append(ref_state)
        return np.hstack(pre_list), cur_vel_array, np.hstack(ref_list), np.array(gear_list) * ref_speed


def pack_paths(paths):
    """B paths (lists of (4,1) arrays) -> the flat layout of nb_ipath_set_paths: points (P,4) float64, per-env curve CSR."""
    pts, curve_begin, env_curve_begin = [], [0], [0]
    for path in paths:
        o = OracleInitialPath(1, 0.1, "diff")
        o.set_initial_path([p.copy() for p in path])
        for curve in o.curve_list:
            pts.extend(p.reshape(4) for p in curve)
            curve_begin.append(len(pts))
        env_curve_begin.append(len(curve_begin) - 1)
    return np.array(pts, dtype=np.float64).reshape(-1, 4), np.array(curve_begin, np.int32), np.array(env_curve_begin, np.int32)
