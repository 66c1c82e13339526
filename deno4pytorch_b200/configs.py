"""Shared layer configuration: the activation table every layer file star-imports
(same keys and module types as /root/reference/Models/configs.py:17-19; ``None`` selects SiLU)."""
import torch.nn as nn

activation_dict = {
    "gelu": nn.GELU(),
    "silu": nn.SiLU(),
    "relu": nn.ReLU(),
    "tanh": nn.Tanh(),
    "leakyrelu": nn.LeakyReLU(),
    None: nn.SiLU(),
}


def default(value, d):
    """``value`` unless it is None."""
    return value if value is not None else d


def activation_kernel_id(module) -> str:
    """Name of the fused-epilogue activation equivalent to ``module``, or '' when the kernel has
    none (the layer then runs the kernel with an identity epilogue and applies ``module`` after).
    Callers may replace ``layer.activation`` at any time (Transformers.py:403 swaps in Identity),
    so this is evaluated per forward call."""
    if isinstance(module, nn.GELU):
        return "gelu" if getattr(module, "approximate", "none") == "none" else ""
    if isinstance(module, nn.SiLU):
        return "silu"
    if isinstance(module, nn.ReLU):
        return "relu"
    if isinstance(module, nn.Tanh):
        return "tanh"
    if isinstance(module, nn.LeakyReLU):
        return "leakyrelu" if module.negative_slope == 0.01 else ""
    if isinstance(module, nn.Identity) or type(module).__name__ == "Identity":
        return "identity"
    return ""
