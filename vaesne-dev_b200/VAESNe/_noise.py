"""Reparameterisation noise: drawn exactly as torch.distributions does (so a seeded global RNG gives
the reference's numbers), or injected by tests for bit-identical parity runs."""
from collections import deque

import torch
import torch.distributions as dist

_injected = deque()

FAMILIES = {dist.Laplace: "laplace", dist.Normal: "normal"}


def family_of(cls) -> str:
    try:
        return FAMILIES[cls]
    except KeyError:
        raise NotImplementedError(
            f"VAESNe-B200 fused path supports Laplace and Normal distributions, got {cls!r}") from None


def inject(noises) -> None:
    """Queue explicit noise tensors; each rsample on the fused path consumes one (FIFO)."""
    _injected.extend(noises)


def clear() -> None:
    _injected.clear()


def draw(family: str, shape, like: torch.Tensor) -> torch.Tensor:
    if _injected:
        n = _injected.popleft()
        if tuple(n.shape) != tuple(shape):
            raise ValueError(f"injected noise has shape {tuple(n.shape)}, expected {tuple(shape)}")
        return n.to(device=like.device, dtype=torch.float32).contiguous()
    if family == "laplace":      # torch/distributions/laplace.py:73-85
        fi = torch.finfo(torch.float32)
        return torch.empty(tuple(shape), dtype=torch.float32, device=like.device).uniform_(fi.eps - 1, 1)
    return torch.empty(tuple(shape), dtype=torch.float32, device=like.device).normal_()
