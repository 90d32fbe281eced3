"""Epoch driver — drop-in for ``training_step`` of the reference's ``VAESNe/training_util.py:17-52``.

Same contract (train mode, per batch: zero_grad, move to the network's device, loss = -loss_fn,
backward, optimizer.step, return the mean loss) with two differences that do not change results:
the per-step ``loss.cpu().item()`` host synchronisation is replaced by an on-device accumulation
that is read back once per epoch, and batches are moved with non-blocking copies.

``cuda_graph=True`` (or ``VAESNE_CUDA_GRAPH=1``) additionally captures the whole step — forward, backward,
fused AdamW — into one CUDA graph per batch shape and replays it: the reference's own batch size (16) is
launch-bound (~470 kernels per step), and every kernel on the path keeps its step counter and dropout seed
on the device precisely so that a replay advances them."""
import gc
import math
import os

import torch

from .losses import elbo


def safelog10(x):
    return math.log10(max(1e-10, x))


def _to_device(batch, device, multimodal):
    if multimodal:
        return [tuple(t.to(device, non_blocking=True) for t in modality) for modality in batch]
    return tuple(t.to(device, non_blocking=True) for t in batch)


def _leaves(batch, multimodal):
    return [t for modality in batch for t in modality] if multimodal else list(batch)


class GraphedStep:
    """One training step (zero_grad, loss = -loss_fn, backward, optimizer.step) as a replayable CUDA graph.

    Per batch signature (shapes + dtypes) the first ``WARM`` batches run eagerly — they are ordinary steps and
    also perform every one-time host initialisation (kernel attributes, TMA encoder, seed cell) — and the next
    one is captured with its inputs living in static buffers; later batches are copied into those buffers and
    the graph is replayed.  Everything the step allocates comes from the graph's private pool, so all device
    pointers baked into the kernel arguments (and the TMA descriptors) stay valid.  A step that cannot be
    captured (a loss with a host synchronisation, a collective in flight) falls back to eager for good."""
    WARM = 2

    def __init__(self, network, optimizer, loss_fn, multimodal):
        self.network, self.optimizer, self.loss_fn, self.multimodal = network, optimizer, loss_fn, multimodal
        self.entries = {}
        self.replayed_launches = 0        # kernels of this library executed through graph replays (bench accounting)

    def _eager(self, x, device):
        self.optimizer.zero_grad()
        x = _to_device(x, device, self.multimodal)
        loss = -self.loss_fn(self.network, x)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def _capture(self, x, device):
        static = _to_device(x, device, self.multimodal)
        static = [tuple(t.clone() for t in m) for m in static] if self.multimodal else tuple(t.clone() for t in static)
        self.optimizer.zero_grad(set_to_none=True)
        # nothing of an earlier (eager) iteration's autograd graph may survive into the capture: its AccumulateGrad nodes live
        # on another stream and the engine would synchronise with it (illegal while capturing).  The models park the last
        # posterior parameters — tensors with a grad_fn — on themselves (base_vae.py:24-30 `_qz_x_params`); drop those.
        for mod in self.network.modules():
            if getattr(mod, "_qz_x_params", None) is not None:
                mod._qz_x_params = None
        gc.collect()
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize(device)
        from . import _native
        l0 = _native.launch_count()
        with torch.cuda.graph(graph):
            loss = -self.loss_fn(self.network, static)
            loss.backward()
            self.optimizer.step()
            out = loss.detach()
        return dict(graph=graph, static=static, out=out, launches=_native.launch_count() - l0)

    def __call__(self, x, device):
        from . import parallel
        leaves = _leaves(x, self.multimodal)
        # learning rate / betas / eps / weight decay enter the captured kernels as host scalars: a scheduler or a param_group
        # edit therefore selects (captures) another graph instead of silently replaying the old values
        hyper = tuple((float(g.get("lr", 0.0)), float(g.get("weight_decay", 0.0)), tuple(g.get("betas", ())), float(g.get("eps", 0.0)))
                      for g in self.optimizer.param_groups)
        sig = tuple((tuple(t.shape), t.dtype) for t in leaves) + (hyper,)
        e = self.entries.setdefault(sig, dict(seen=0))
        # data-parallel runs are captured too: the per-stack bucket all-reduces are NCCL launches on NCCL's own stream, forked
        # from / joined to the capturing stream by events, so they become nodes of the graph (VAESNE_DP_GRAPH=0 opts out)
        if e.get("eager") or device.type != "cuda" or (parallel.enabled() and os.environ.get("VAESNE_DP_GRAPH", "1") in ("", "0")):
            return self._eager(x, device)
        if "graph" not in e:
            if e["seen"] < self.WARM:
                e["seen"] += 1
                return self._eager(x, device)
            try:
                e.update(self._capture(x, device))
            except Exception as err:      # noqa: BLE001 — any capture failure means "this step is not capturable"
                e["eager"] = True
                torch.cuda.synchronize(device)
                import warnings
                warnings.warn(f"VAESNe: CUDA-graph capture of the training step failed ({err!r}); running eagerly")
                return self._eager(x, device)
        else:
            for s, t in zip(_leaves(e["static"], self.multimodal), leaves):
                s.copy_(t, non_blocking=True)
        e["graph"].replay()
        self.replayed_launches += e["launches"]
        return e["out"].clone()


def _graphed_step(network, optimizer, loss_fn, multimodal):
    """The GraphedStep of (network, loss_fn) — kept on the optimizer, so it lives exactly as long as the training run."""
    table = optimizer.__dict__.setdefault("_vaesne_graphed", {})
    key = (id(network), loss_fn, bool(multimodal))
    g = table.get(key)
    if g is None or g.network is not network:
        g = table[key] = GraphedStep(network, optimizer, loss_fn, multimodal)
    return g


def training_step(network, optimizer, data_loader, loss_fn=elbo, multimodal=False, release_memory=False, cuda_graph=None):
    """Train for one epoch; returns the average of the per-batch losses (a Python float)."""
    network.train()
    device = next(network.parameters()).device
    if cuda_graph is None:
        cuda_graph = os.environ.get("VAESNE_CUDA_GRAPH", "0") not in ("", "0")
    graphed = _graphed_step(network, optimizer, loss_fn, multimodal) if (cuda_graph and device.type == "cuda") else None
    losses = []
    for x in data_loader:
        if graphed is not None:
            losses.append(graphed(x, device))
            continue
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
