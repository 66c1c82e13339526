from Models.fno.FNOs import *  # noqa: F401,F403
from Models.fno.FNOs import FNO1d, FNO2d, FNO3d  # noqa: F401
