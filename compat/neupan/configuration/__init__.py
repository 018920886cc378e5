from neupan_b200.configuration import *  # noqa: F401,F403
from neupan_b200 import configuration as _c
device, time_print, tensor_dtype = _c.device, _c.time_print, _c.tensor_dtype
