"""Drop-in for /root/reference/Models/configs.py (the names the FNO path star-imports)."""
import math  # noqa: F401

import numpy as np  # noqa: F401
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F  # noqa: F401

from deno4pytorch_b200.configs import activation_dict, default  # noqa: F401
