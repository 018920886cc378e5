"""Host-side logic of the boundary: geometry, decimation, robot attributes (no GPU)."""
import numpy as np

from helpers import GOLDEN
from neupan_b200.robot import robot
from neupan_b200.util import decimation_indices, downsample_decimation, gen_inequal_from_vertex


def test_gen_inequal_matches_reference_golden():
    z = np.load(f"{GOLDEN}/ref_misc.npz")
    for k in ("rect_diff", "polygon_cw", "pentagon"):
        G, h = gen_inequal_from_vertex(z[f"poly_{k}_v"])
        assert np.array_equal(G, z[f"poly_{k}_G"]) and np.array_equal(h, z[f"poly_{k}_h"])
    assert gen_inequal_from_vertex(np.array([[0, 2, 1, 2, 0], [0, 0, 1, 2, 2]], float)) == (None, None)  # non-convex


def test_decimation_matches_reference_golden():
    z = np.load(f"{GOLDEN}/ref_misc.npz")
    assert np.array_equal(decimation_indices(500, 100), z["decim_idx_500_100"])
    assert np.array_equal(downsample_decimation(np.arange(1080)[None, :], 100)[0], z["decim_idx_1080_100"])
    a = np.arange(10)[None, :]
    assert downsample_decimation(a, 20) is a


def test_robot_matches_shipped_geometry():
    # G, h of the three shipped robots (headers of example/model/*/results.txt)
    r = robot(10, 0.1, kinematics="diff", length=1.6, width=2.0)
    assert np.allclose(r.G, [[0, -1.6], [2, 0], [0, 1.6], [-2, 0]]) and np.allclose(r.h.ravel(), 1.6)
    r = robot(10, 0.1, kinematics="acker", length=4.6, width=1.6, wheelbase=3, max_speed=[8, 2], max_acce=[8, 1])
    assert np.allclose(r.G, [[0, -4.6], [1.6, 0], [0, 4.6], [-1.6, 0]]) and np.allclose(r.h.ravel(), [3.68, 6.08, 3.68, 1.28])
    assert r.max_speed[1, 0] == 1.57 and np.allclose(r.acce_bound.ravel(), [0.8, 0.1])
    r = robot(10, 0.1, kinematics="omni", vertices=[[-0.8, -1.0], [-1.8, 1.0], [1.8, 1.0], [0.8, -1.0]])
    assert np.allclose(r.G, [[0, -1.6], [2, -1], [0, 3.6], [-2, -1]]) and np.allclose(r.h.ravel(), [1.6, 2.6, 3.6, 2.6])
    try:
        robot(10, 0.1)
        assert False
    except ValueError:
        pass
