"""The ALGORITHM of csrc/ipath_kernel.cuh, transcribed statement by statement to Python (flat path array, alias bookkeeping with
write-through and read-back, numpy's fused two-element dot), replayed on the reference's golden runs: it must reproduce the
reference class bit for bit, including the path the reference leaves behind.  This is the CPU-side proof that the kernel's way
of emulating numpy's view semantics is equivalent to the reference; the GPU tests (tests/test_ipath.py) then only have to show
that the CUDA code computes what this transcription computes (up to FP64 cos/sin/tan ulps)."""
import importlib.util
import math
import os
from fractions import Fraction

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_ipath", os.path.join(HERE, "golden", "make_golden_ipath.py"))
mgi = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mgi)
GOLD = np.load(os.path.join(HERE, "golden", "ref_ipath.npz"))
PI = math.pi


def fma(a, b, c):  # exact product and sum, one rounding
    return float(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def dot2(x0, x1, y0, y1):  # ipath_kernel.cuh: dot2
    return fma(x1, y1, x0 * y0)


def wrap(r):
    while r > PI:
        r = r - 2 * PI
    while r < -PI:
        r = r + 2 * PI
    return r


class KernelAlgorithm:
    def __init__(self, T, dt, kin, L, loop, path):
        self.T, self.dt, self.kin, self.L, self.loop = T, dt, kin, L, loop
        self.P = np.hstack(path).T.copy()  # (P,4), mutable
        n, acc = len(self.P), 0.0
        for i in range(n - 1):  # InitialPathBatch.set_initial_paths
            acc += math.hypot(self.P[i + 1, 0] - self.P[i, 0], self.P[i + 1, 1] - self.P[i, 1])
        self.interval = acc / (n - 1)
        self.cb = [0] + [i for i in range(1, n) if self.P[i, 3] != self.P[i - 1, 3]] + [n]
        self.ci = self.pi = 0
        self.ind_range, self.ait, self.at, self.ct = 10, 1, 0.1, 0.1

    def step(self, state, vel, ref_speed):
        T, P = self.T, self.P
        sx, sy, sth = (float(v) for v in state)
        ci, pi = self.ci, self.pi
        p0 = self.cb[ci]
        ln = self.cb[ci + 1] - p0
        md = math.inf
        for i in range(max(pi, 0), min(pi + self.ind_range, ln)):  # closest_point
            dx, dy = sx - P[p0 + i, 0], sy - P[p0 + i, 1]
            d = math.sqrt(dx * dx + dy * dy)
            if d < md:
                md, pi = d, i
                if d < self.ct:
                    break
        dx, dy = sx - P[p0 + ln - 1, 0], sy - P[p0 + ln - 1, 1]
        ret = False
        if math.sqrt(dot2(dx, dy, dx, dy)) < self.at and pi >= ln - self.ait - 2:
            if ci + 1 >= len(self.cb) - 1:
                if self.loop:
                    ci = pi = 0
                else:
                    ret = True
            else:
                ci, pi = ci + 1, 0
            p0 = self.cb[ci]
            ln = self.cb[ci + 1] - p0
        self.ci, self.pi = ci, pi
        if ret:
            return True, None
        rx, ry, rth = (float(v) for v in P[p0 + pi, 0:3])
        alias, ref_index = -1, pi
        px, py, pth = sx, sy, sth
        gear0, fwd = float(P[p0 + pi, 3]), ref_speed * self.dt
        ns, rs, ru, alias_of = np.zeros((3, T + 1)), np.zeros((3, T + 1)), np.zeros(T), [-1] * T
        ns[:, 0], rs[:, 0] = (px, py, pth), (rx, ry, rth)
        for t in range(T):
            v, w = float(vel[0, t]), float(vel[1, t])
            if self.kin == "acker":
                px, py, pth = px + (v * math.cos(pth)) * self.dt, py + (v * math.sin(pth)) * self.dt, pth + (v * math.tan(w) / self.L) * self.dt
            elif self.kin == "diff":
                px, py, pth = px + (v * math.cos(pth)) * self.dt, py + (v * math.sin(pth)) * self.dt, pth + w * self.dt
            else:
                px, py, pth = px + self.dt * (v * math.cos(w)), py + self.dt * (v * math.sin(w)), pth + self.dt * 0.0
            ns[:, t + 1] = (px, py, pth)
            gear = gear0
            if fwd >= self.interval:
                ref_index = ref_index + int(fwd / self.interval)
                if ref_index > ln - 1:
                    ref_index, gear = ln - 1, 0.0
                alias = ref_index
            else:
                cx, cy = (P[p0 + alias, 0], P[p0 + alias, 1]) if alias >= 0 else (rx, ry)
                while True:
                    if ref_index > ln - 2:
                        P[p0 + ln - 1, 2] = wrap(P[p0 + ln - 1, 2])
                        alias = ln - 1
                        break
                    ax, ay, bx, by = P[p0 + ref_index, 0], P[p0 + ref_index, 1], P[p0 + ref_index + 1, 0], P[p0 + ref_index + 1, 1]
                    dx, dy, hit = bx - ax, by - ay, False
                    if math.sqrt(dot2(dx, dy, dx, dy)) != 0.0:
                        fx, fy = ax - cx, ay - cy
                        a, bq, c = dot2(dx, dy, dx, dy), 2.0 * dot2(fx, fy, dx, dy), dot2(fx, fy, fx, fy) - fwd * fwd
                        disc = bq * bq - (4.0 * a) * c
                        if not disc < 0:
                            t2 = (-bq + math.sqrt(disc)) / (2.0 * a)
                            if 0 <= t2 <= 1:
                                rx, ry = ax + t2 * dx, ay + t2 * dy
                                diff = wrap(P[p0 + ref_index + 1, 2] - P[p0 + ref_index, 2])
                                rth, alias, hit = wrap(P[p0 + ref_index, 2] + diff / 2), -1, True
                    if hit:
                        break
                    ref_index += 1
                if ref_index > ln - 1:
                    gear = 0.0
            if alias >= 0:
                P[p0 + alias, 2] = pth + wrap(P[p0 + alias, 2] - pth)
            else:
                rth = pth + wrap(rth - pth)
                rs[:, t + 1] = (rx, ry, rth)
            alias_of[t], ru[t] = alias, gear * ref_speed
        for t in range(T):
            if alias_of[t] >= 0:
                rs[:, t + 1] = P[p0 + alias_of[t], 0:3]
        return False, (ns, rs, ru)


@pytest.mark.parametrize("name", list(mgi.SCENARIOS))
def test_kernel_algorithm_reproduces_the_reference_bit_for_bit(name):
    kin, L, loop, step, split, curve, n = mgi.SCENARIOS[name]
    g = {k.split(".", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith(name + ".")}
    alg = KernelAlgorithm(mgi.T, mgi.DT, kin, L, loop, mgi.make_path(n, step, split, curve))
    for k in range(len(g["arrived"])):
        arrived, out = alg.step(g["states"][k], g["vel"][k], mgi.REF_SPEED)
        assert arrived == bool(g["arrived"][k]) and alg.pi == g["point_index"][k] and alg.ci == g["curve_index"][k]
        if arrived:
            break
        assert np.array_equal(out[0], g["nom_s"][k]) and np.array_equal(out[1], g["ref_s"][k]) and np.array_equal(out[2], g["ref_us"][k])
    assert np.array_equal(alg.P, g["final_path"])
