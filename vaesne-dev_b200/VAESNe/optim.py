"""Fused flat AdamW for the VAESNe models.

``torch.optim.AdamW`` (what the reference scripts construct, e.g. cannon/test_photospectra.py:135)
launches a handful of foreach kernels over ~350 small tensors.  Here all trainable parameters of a
group are re-homed as views of one flat fp32 buffer, gradients (which the stacks already emit as a
few flat buckets) are gathered with one strided copy per bucket, and a single kernel applies the
decoupled-weight-decay Adam update.  The step counter lives on the device so the whole training step
can be captured in a CUDA graph.  Same update rule and hyper-parameter meaning as torch's AdamW."""
from __future__ import annotations

import torch

from . import _ops as P
from . import parallel


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_average=False):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.grad_average = grad_average
        self._flat = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.requires_grad]
            if not ps:
                self._flat.append(None)
                continue
            dev = ps[0].device
            sizes = [p.numel() for p in ps]
            flat_p = torch.empty(sum(sizes), device=dev, dtype=torch.float32)
            offs, off = [], 0
            for p, n in zip(ps, sizes):
                if p.dtype != torch.float32 or p.device != dev:
                    raise TypeError("FusedAdamW needs fp32 parameters on one device")
                flat_p[off:off + n].copy_(p.data.reshape(-1))
                p.data = flat_p[off:off + n].view(p.shape)
                offs.append(off)
                off += n
            st = dict(ps=ps, sizes=sizes, offs=offs, p=flat_p, g=torch.zeros_like(flat_p), m=torch.zeros_like(flat_p),
                      v=torch.zeros_like(flat_p), step=torch.zeros(1, dtype=torch.int32, device=dev))
            self._flat.append(st)

    @torch.no_grad()
    def _gather(self, st):
        """Copy gradients into the flat buffer, one strided copy per run of parameters whose grads already sit
        back-to-back in one bucket.  Returns the [start, end) element ranges that received a gradient."""
        ps, sizes, offs, g = st["ps"], st["sizes"], st["offs"], st["g"]
        ranges = []
        i, n = 0, len(ps)
        while i < n:
            gi = ps[i].grad
            if gi is None:
                i += 1
                continue
            if gi.dtype != torch.float32 or not gi.is_contiguous():
                gi = gi.float().contiguous()
            start_ptr, start_off, run = gi.data_ptr(), offs[i], sizes[i]
            keep = [gi]
            j = i + 1
            while j < n:
                gj = ps[j].grad
                if gj is None or gj.dtype != torch.float32 or not gj.is_contiguous() or gj.data_ptr() != start_ptr + 4 * run:
                    break
                keep.append(gj)
                run += sizes[j]
                j += 1
            src = keep[0]
            P.N.check(P.N.lib().vaesne_copy3d(start_ptr, 0, 0, g.data_ptr() + 4 * start_off, 0, 0, 1, 1, run, 0, P.N.stream_of(src)))
            if ranges and ranges[-1][1] == start_off:
                ranges[-1][1] = start_off + run
            else:
                ranges.append([start_off, start_off + run])
            i = j
        return ranges

    # ---- checkpointing: the moments and the step counter live in the flat buffers; they are exposed in (and restored from)
    # torch.optim.AdamW's own state layout, so optimizer.state_dict() files are interchangeable with the reference's optimiser
    def _mirror_state(self):
        for st in self._flat:
            if st is None:
                continue
            for p, off, n in zip(st["ps"], st["offs"], st["sizes"]):
                self.state[p] = {"step": st["step"].to(torch.float32).reshape(()).clone(),
                                 "exp_avg": st["m"][off:off + n].view(p.shape), "exp_avg_sq": st["v"][off:off + n].view(p.shape)}

    def state_dict(self):
        self._mirror_state()
        return super().state_dict()

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for st in self._flat:
            if st is None:
                continue
            steps = []
            for p, off, n in zip(st["ps"], st["offs"], st["sizes"]):
                s = self.state.get(p)
                if not s:
                    continue
                st["m"][off:off + n].copy_(s["exp_avg"].reshape(-1))
                st["v"][off:off + n].copy_(s["exp_avg_sq"].reshape(-1))
                steps.append(int(float(s["step"])))
            if steps:
                if len(set(steps)) != 1:
                    raise ValueError("FusedAdamW keeps one step counter per parameter group; the loaded state has several")
                st["step"].fill_(steps[0])
        self._mirror_state()

    def hyper_signature(self):
        """What a captured CUDA graph of step() has baked in as host scalars (training_util re-captures when it changes)."""
        return tuple((float(g["lr"]), tuple(float(b) for b in g["betas"]), float(g["eps"]), float(g["weight_decay"])) for g in self.param_groups)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        parallel.wait_all()
        scale = 1.0 / parallel.world_size() if (self.grad_average and parallel.enabled()) else 1.0
        for group, st in zip(self.param_groups, self._flat):
            if st is None:
                continue
            ranges = self._gather(st)
            if not ranges:
                continue
            P.step_advance(st["step"], None)
            b1, b2 = group["betas"]
            for a, b in ranges:
                P.adamw_flat(st["p"][a:b], st["g"][a:b], st["m"][a:b], st["v"][a:b], group["lr"], b1, b2, group["eps"],
                             group["weight_decay"], st["step"], scale)
        return loss
