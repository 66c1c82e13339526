"""B200-native FNO spectral-convolution path (drop-in for DENO4pytorch's Models/fno).

Heavy submodules are imported lazily so that ``deno4pytorch_b200._capi`` (pure ctypes) can be
used without touching torch or the CUDA library.
"""
__version__ = "0.1.0"

_LAZY = {
    "SpectralConv1d": "spectral_layers", "SpectralConv2d": "spectral_layers", "SpectralConv3d": "spectral_layers",
    "FNO1d": "FNOs", "FNO2d": "FNOs", "FNO3d": "FNOs",
    "AdaptiveFourier1d": "spectral_layers", "AdaptiveFourier2d": "spectral_layers", "AdaptiveFourier3d": "spectral_layers",
    "spectral_conv": "functional", "afno_filter": "functional",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod = importlib.import_module(f"{__name__}.{_LAZY[name]}")
        return getattr(mod, name)
    raise AttributeError(name)
