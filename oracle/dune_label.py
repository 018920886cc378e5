"""Labels of the DUNE training set in closed form (TEST INFRASTRUCTURE; groundwork for SURVEY 8f row 4).

``DUNETrain.prob_solve`` (neupan/blocks/dune_train.py:83-98, 134-140) solves, per sampled point p, the cone program (10)

    max_mu  mu' (G p - h)    s.t.  ||G' mu|| <= 1,  mu >= 0

with cvxpy/ECOS and stores (p, mu*, value) as (input, label, distance) (process_data / generate_data_set, :100-132).  The
program is the dual of the distance from p to the convex polygon {x : G x <= h}, so its solution is available without a
solver: value = dist(p, polygon); the maximiser is supported on the edge(s) whose outward normal cone contains p - x*, x* the
closest point of the polygon:
    x* inside an edge e        ->  mu_e = 1 / ||G_e||, the rest 0
    x* a vertex of edges e, f  ->  (mu_e, mu_f) >= 0 solve  G_e' mu_e + G_f' mu_f = (p - x*) / ||p - x*||
    p inside the polygon       ->  mu = 0, value 0
``certificate`` checks a label without reference to how it was computed: dual feasibility and zero duality gap against the
primal distance obtained by plain geometry.  tests/test_oracle_dune_label.py pins it to the polygons of the shipped robots.
"""
import numpy as np


def polygon_vertices(G, h):
    """Vertices of {x : G x <= h} in the cyclic order of the rows (adjacent rows of gen_inequal_from_vertex share a vertex)."""
    G, h = np.asarray(G, np.float64), np.asarray(h, np.float64).reshape(-1)
    E = G.shape[0]
    return np.array([np.linalg.solve(np.vstack([G[e], G[(e + 1) % E]]), np.array([h[e], h[(e + 1) % E]])) for e in range(E)])  # vertex e = rows e, e+1


def label(G, h, p):
    """(mu* (E,), distance) for one point p (2,)."""
    G, h, p = np.asarray(G, np.float64), np.asarray(h, np.float64).reshape(-1), np.asarray(p, np.float64).reshape(2)
    E = G.shape[0]
    r = G @ p - h
    mu = np.zeros(E)
    if np.all(r <= 0):
        return mu, 0.0
    V = polygon_vertices(G, h)  # edge e+1 runs from vertex e to vertex e+1 (it is the row shared by both)
    best = (np.inf, None)
    for e in range(E):
        a, b = V[(e - 1) % E], V[e]  # the two vertices on row e
        d = b - a
        t = np.clip((p - a) @ d / (d @ d), 0.0, 1.0)
        x = a + t * d
        dist = np.linalg.norm(p - x)
        if dist < best[0]:
            best = (dist, (e, t, x))
    dist, (e, t, x) = best
    if 0.0 < t < 1.0:  # interior of edge e
        mu[e] = 1.0 / np.linalg.norm(G[e])
        return mu, float(r[e] * mu[e])
    f = (e - 1) % E if t == 0.0 else (e + 1) % E  # the other row through the vertex
    nvec = (p - x) / dist
    sol = np.linalg.solve(np.vstack([G[e], G[f]]).T, nvec)
    mu[e], mu[f] = max(sol[0], 0.0), max(sol[1], 0.0)
    return mu, float(mu @ r)


def labels(G, h, points):
    out = [label(G, h, p) for p in np.asarray(points, np.float64).reshape(-1, 2)]
    return np.array([o[0] for o in out]), np.array([o[1] for o in out])


def primal_distance(G, h, p):
    """dist(p, polygon) by geometry alone: 0 inside, else the smallest point-to-segment distance over the boundary."""
    G, h, p = np.asarray(G, np.float64), np.asarray(h, np.float64).reshape(-1), np.asarray(p, np.float64).reshape(2)
    if np.all(G @ p - h <= 0):
        return 0.0
    V = polygon_vertices(G, h)
    best = np.inf
    for e in range(len(V)):
        a, b = V[e], V[(e + 1) % len(V)]
        d = b - a
        t = np.clip((p - a) @ d / (d @ d), 0.0, 1.0)
        best = min(best, np.linalg.norm(p - (a + t * d)))
    return float(best)


def certificate(G, h, p, mu, value):
    """(dual infeasibility, |duality gap|): both ~0 iff (mu, value) is optimal for program (10)."""
    G, h = np.asarray(G, np.float64), np.asarray(h, np.float64).reshape(-1)
    infeas = max(0.0, np.linalg.norm(G.T @ mu) - 1.0, float(-mu.min()))
    return infeas, abs(float(mu @ (G @ np.asarray(p, np.float64).reshape(2) - h)) - primal_distance(G, h, p)) + abs(value - float(mu @ (G @ np.asarray(p, np.float64).reshape(2) - h)))
