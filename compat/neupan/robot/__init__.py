from neupan_b200.robot import robot  # noqa: F401
