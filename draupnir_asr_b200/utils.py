"""Host helpers the hot path shares with the reference's utils (S/utils.py)."""
from __future__ import annotations

import torch


def squeeze_tensor(required_ndims: int, tensor: torch.Tensor) -> torch.Tensor:
    """Drop size-1 dimensions until `required_ndims` remain (S/utils.py:1874-1902); used by gp_prior on the
    `[z,1]`-shaped PyroParams of the variational guide."""
    while tensor.dim() > required_ndims:
        if tensor.shape[0] == 1:
            tensor = tensor.squeeze(0)
        elif 1 in tensor.shape:
            tensor = tensor.squeeze(list(tensor.shape).index(1))
        else:
            break
    return tensor


def require_cuda(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"draupnir_asr_b200 runs on CUDA (sm_100a) only; got device {dev!r}. There is no CPU path — "
                           "use the reference implementation for CPU runs.")
    return dev
