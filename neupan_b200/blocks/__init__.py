from .obs_point_net import ObsPointNet
from .dune import DUNE
from .nrmp import NRMP
from .pan import PAN
from .initial_path import InitialPath

__all__ = ["ObsPointNet", "DUNE", "NRMP", "PAN", "InitialPath"]
