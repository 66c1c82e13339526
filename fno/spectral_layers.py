from Models.fno.spectral_layers import *  # noqa: F401,F403
