from .obs_point_net import ObsPointNet
from .dune import DUNE
from .nrmp import NRMP
from .pan import PAN

__all__ = ["ObsPointNet", "DUNE", "NRMP", "PAN"]
