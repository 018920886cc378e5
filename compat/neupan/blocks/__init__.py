from neupan_b200.blocks import DUNE, NRMP, PAN, InitialPath, ObsPointNet  # noqa: F401
