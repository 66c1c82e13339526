"""``Models.fno.FNOs`` / ``fno.FNOs`` -> deno4pytorch_b200.FNOs."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from deno4pytorch_b200.FNOs import *  # noqa: F401,F403,E402
from deno4pytorch_b200.FNOs import FNO1d, FNO2d, FNO3d  # noqa: F401,E402
