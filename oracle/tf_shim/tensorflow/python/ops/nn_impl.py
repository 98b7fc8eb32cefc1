import torch


def l2_normalize(x, axis=None, epsilon=1e-12):
    """[TF] x * rsqrt(maximum(reduce_sum(square(x), axis, keepdims=True), epsilon))."""
    ss = (x * x).sum(dim=tuple(axis) if isinstance(axis, (list, tuple)) else axis, keepdim=True)
    return x * torch.rsqrt(torch.clamp(ss, min=epsilon))
