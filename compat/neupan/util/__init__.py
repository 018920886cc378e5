from neupan_b200.util import *  # noqa: F401,F403
