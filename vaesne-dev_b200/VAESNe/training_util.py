"""Epoch driver — drop-in for ``training_step`` of the reference's ``VAESNe/training_util.py:17-52``.

Same contract (train mode, per batch: zero_grad, move to the network's device, loss = -loss_fn,
backward, optimizer.step, return the mean loss) with two differences that do not change results:
the per-step ``loss.cpu().item()`` host synchronisation is replaced by an on-device accumulation
that is read back once per epoch, and batches are moved with non-blocking copies."""
import gc
import math

import torch

from .losses import elbo


def safelog10(x):
    return math.log10(max(1e-10, x))


def _to_device(batch, device, multimodal):
    if multimodal:
        return [tuple(t.to(device, non_blocking=True) for t in modality) for modality in batch]
    return tuple(t.to(device, non_blocking=True) for t in batch)


def training_step(network, optimizer, data_loader, loss_fn=elbo, multimodal=False, release_memory=False):
    """Train for one epoch; returns the average of the per-batch losses (a Python float)."""
    network.train()
    device = next(network.parameters()).device
    losses = []
    for x in data_loader:
        optimizer.zero_grad()
        x = _to_device(x, device, multimodal)
        loss = -loss_fn(network, x)
        loss.backward()
        optimizer.step()
        losses.append(loss.detach())
        if release_memory:
            del x
            gc.collect()
            if torch.cuda.is_available():
                torch.cuda.empty_cache()
    if not losses:
        return float("nan")
    return float(torch.stack(losses).double().mean().cpu().item())
