"""Reeds-Shepp curves for the host-side curve generator (``curve_style: 'reeds'``; the reference gets them from the
third-party gctl package, neupan/blocks/initial_path.py:22,330-332, which is absent here).

Candidate words of the classic families -- C|S|C, C|C|C, CC|CC, C|CSC, C|CSC|C, each under the time-flip / reflection
symmetries -- are generated from the closed-form expressions of Reeds & Shepp (1990) in the goal-centred normal form
(unit turning radius, start at the origin heading +x).  Every candidate is then VERIFIED by integrating it and only words
that actually reach the goal pose are kept; the shortest verified word is returned.  A forward-only word (Dubins) is always
among the candidates, so a path always exists.
"""
from __future__ import annotations

import math
from math import atan2, cos, pi, sin, sqrt


def _mod2pi(x):
    v = math.fmod(x, 2 * pi)
    if v < -pi:
        v += 2 * pi
    elif v > pi:
        v -= 2 * pi
    return v


def _polar(x, y):
    return math.hypot(x, y), atan2(y, x)


# ---- base words in normal form: return (t, u, v) or None --------------------------------------------
def _lsl(x, y, phi):
    u, t = _polar(x - sin(phi), y - 1 + cos(phi))
    if t >= 0:
        v = _mod2pi(phi - t)
        if v >= 0:
            return t, u, v
    return None


def _lsr(x, y, phi):
    u1, t1 = _polar(x + sin(phi), y - 1 - cos(phi))
    if u1 * u1 >= 4:
        u = sqrt(u1 * u1 - 4)
        t = _mod2pi(t1 + atan2(2, u))
        v = _mod2pi(t - phi)
        if t >= 0 and v >= 0:
            return t, u, v
    return None


def _lrl(x, y, phi):
    u1, t1 = _polar(x - sin(phi), y - 1 + cos(phi))
    if u1 <= 4:
        u = -2 * math.asin(0.25 * u1)
        t = _mod2pi(t1 + 0.5 * u + pi)
        v = _mod2pi(phi - t + u)
        if t >= 0 and u <= 0:
            return t, u, v
    return None


def _tau_omega(u, v, xi, eta, phi):
    delta = _mod2pi(u - v)
    A, B = sin(u) - sin(delta), cos(u) - cos(delta) - 1
    t1 = atan2(eta * A - xi * B, xi * A + eta * B)
    t2 = 2 * (cos(delta) - cos(v) - cos(u)) + 3
    tau = _mod2pi(t1 + pi) if t2 < 0 else _mod2pi(t1)
    return tau, _mod2pi(tau - u + v - phi)


def _lrlrn(x, y, phi):
    xi, eta = x + sin(phi), y - 1 - cos(phi)
    rho = 0.25 * (2 + sqrt(xi * xi + eta * eta))
    if rho <= 1:
        u = math.acos(rho)
        t, v = _tau_omega(u, -u, xi, eta, phi)
        if t >= 0 and v <= 0:
            return t, u, v
    return None


def _lrlrp(x, y, phi):
    xi, eta = x + sin(phi), y - 1 - cos(phi)
    rho = (20 - xi * xi - eta * eta) / 16
    if 0 <= rho <= 1:
        u = -math.acos(rho)
        if u >= -0.5 * pi:
            t, v = _tau_omega(u, u, xi, eta, phi)
            if t >= 0 and v >= 0:
                return t, u, v
    return None


def _lrsl(x, y, phi):
    xi, eta = x - sin(phi), y - 1 + cos(phi)
    rho, theta = _polar(xi, eta)
    if rho >= 2:
        r = sqrt(rho * rho - 4)
        u = 2 - r
        t = _mod2pi(theta + atan2(r, -2))
        v = _mod2pi(phi - 0.5 * pi - t)
        if t >= 0 and u <= 0 and v <= 0:
            return t, u, v
    return None


def _lrsr(x, y, phi):
    xi, eta = x + sin(phi), y - 1 - cos(phi)
    rho, theta = _polar(-eta, xi)
    if rho >= 2:
        t, u = theta, 2 - rho
        v = _mod2pi(t + 0.5 * pi - phi)
        if t >= 0 and u <= 0 and v <= 0:
            return t, u, v
    return None


def _lrslr(x, y, phi):
    xi, eta = x + sin(phi), y - 1 - cos(phi)
    rho, _ = _polar(xi, eta)
    if rho >= 2:
        u = 4 - sqrt(rho * rho - 4)
        if u <= 0:
            t = _mod2pi(atan2((4 - u) * xi - 2 * eta, -2 * xi + (u - 4) * eta))
            v = _mod2pi(t - phi)
            if t >= 0 and v >= 0:
                return t, u, v
    return None


def _candidates(x, y, phi):
    """(types, signed lengths) of every candidate word in normal form."""
    out = []

    def add(types, lengths, flip, reflect):
        if flip:
            lengths = [-l for l in lengths]
        if reflect:
            types = "".join({"L": "R", "R": "L", "S": "S"}[c] for c in types)
        out.append((types, lengths))

    for flip in (False, True):
        for reflect in (False, True):
            xx, yy, pp = (-x if flip else x), (-y if reflect else y), phi
            if flip != reflect:
                pp = -phi
            r = _lsl(xx, yy, pp)
            if r: add("LSL", list(r), flip, reflect)
            r = _lsr(xx, yy, pp)
            if r: add("LSR", list(r), flip, reflect)
            r = _lrl(xx, yy, pp)
            if r: add("LRL", list(r), flip, reflect)
            r = _lrlrn(xx, yy, pp)
            if r: add("LRLR", [r[0], r[1], -r[1], r[2]], flip, reflect)
            r = _lrlrp(xx, yy, pp)
            if r: add("LRLR", [r[0], r[1], r[1], r[2]], flip, reflect)
            r = _lrsl(xx, yy, pp)
            if r: add("LRSL", [r[0], -0.5 * pi, r[1], r[2]], flip, reflect)
            r = _lrsr(xx, yy, pp)
            if r: add("LRSR", [r[0], -0.5 * pi, r[1], r[2]], flip, reflect)
            r = _lrslr(xx, yy, pp)
            if r: add("LRSLR", [r[0], -0.5 * pi, r[1], -0.5 * pi, r[2]], flip, reflect)
            # the same words traversed backwards (goal seen from the start with the roles exchanged)
            xb, yb = xx * cos(pp) + yy * sin(pp), xx * sin(pp) - yy * cos(pp)
            r = _lrl(xb, yb, pp)
            if r: add("LRL", [r[2], r[1], r[0]], flip, reflect)
            r = _lrsl(xb, yb, pp)
            if r: add("LSRL", [r[2], r[1], -0.5 * pi, r[0]], flip, reflect)
            r = _lrsr(xb, yb, pp)
            if r: add("RSRL", [r[2], r[1], -0.5 * pi, r[0]], flip, reflect)
    return out


def integrate(types, lengths, s=None):
    """Pose (x, y, heading) and gear after travelling arc length s (default: the whole word) from the origin, unit radius."""
    x = y = h = 0.0
    gear = 1.0
    rem = sum(abs(l) for l in lengths) if s is None else s
    for c, l in zip(types, lengths):
        if abs(l) == 0.0:
            continue
        step = min(rem, abs(l))
        d = math.copysign(step, l)
        gear = 1.0 if l > 0 else -1.0
        if c == "S":
            x += d * cos(h); y += d * sin(h)
        else:
            sg = 1.0 if c == "L" else -1.0
            x += sg * (sin(h + sg * d) - sin(h)); y += -sg * (cos(h + sg * d) - cos(h))
            h += sg * d
        rem -= step
        if rem <= 0:
            break
    return x, y, h, gear


def shortest_word(x, y, phi, tol=1e-6):
    """Shortest VERIFIED word reaching (x, y, phi) from the origin in normal form: (types, lengths, total length)."""
    best = None
    for types, lengths in _candidates(x, y, phi):
        total = sum(abs(l) for l in lengths)
        if best is not None and total >= best[2]:
            continue
        ex, ey, eh, _ = integrate(types, lengths)
        if math.hypot(ex - x, ey - y) < tol and abs(_mod2pi(eh - phi)) < tol:
            best = (types, lengths, total)
    if best is None:
        raise RuntimeError("no Reeds-Shepp word reached the goal (should not happen: the Dubins words are candidates)")
    return best


def sample_path(a, b, step, r):
    """Points [(x, y, heading, gear)] every `step` along the shortest Reeds-Shepp curve from pose a to pose b, radius r."""
    dx, dy = (b[0] - a[0]) / r, (b[1] - a[1]) / r
    c, s = cos(a[2]), sin(a[2])
    types, lengths, total = shortest_word(c * dx + s * dy, -s * dx + c * dy, _mod2pi(b[2] - a[2]))
    L = total * r
    n = max(int(math.floor(L / step)), 0) if step > 0 else 0
    dists = [k * step for k in range(n + 1)] + ([L] if L - n * step > 1e-9 else [])
    pts = []
    for d in dists:
        x, y, h, gear = integrate(types, lengths, d / r)
        pts.append((a[0] + r * (c * x - s * y), a[1] + r * (s * x + c * y), _mod2pi(a[2] + h), gear))
    return pts
