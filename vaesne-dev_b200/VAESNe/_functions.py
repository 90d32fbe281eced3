"""torch.autograd bridges: one Function per stack, one for the latent step, one per objective.

Parameter gradients of a stack are produced as views of ONE freshly zeroed flat fp32 buffer laid
out in ``named_parameters()`` order — autograd's AccumulateGrad adopts the views as ``.grad`` (no
extra kernels), the fused AdamW and the data-parallel all-reduce then see a few large buffers
instead of ~350 small tensors.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from . import _ops as P
from . import _stacks as S
from . import parallel


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def dev_guard(device):
    """The kernels are enqueued on torch's current stream OF THE TENSORS' DEVICE; a model that lives on a device other than the
    process's current one (device='cuda:1' without torch.cuda.set_device) therefore runs under a device guard."""
    if device.type == "cuda" and device.index is not None and device.index != torch.cuda.current_device():
        return torch.cuda.device(device)
    return _NO_GUARD


def _prep(t: Optional[torch.Tensor], dtype=None):
    if t is None:
        return None
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t if t.is_contiguous() else t.contiguous()


class _StackFn(torch.autograd.Function):
    """forward(runner, names, n_data, *data, *params) -> output tensor.

    `runner(tape, pv, *data)` enqueues the kernels; tensors in `data` that require grad (the latent
    samples fed to a decoder) get their gradient from the tape."""

    @staticmethod
    def forward(ctx, runner: Callable, names: Sequence[str], n_data: int, drop_p: float, *args):
        data, params = args[:n_data], args[n_data:]
        need = any(ctx.needs_input_grad)
        dev = params[0].device
        with dev_guard(dev):
            tape = S.Tape(need, drop_p, dev)
            pv = S.PView(names, [p.detach() for p in params])
            out = runner(tape, pv, *[d.detach() if isinstance(d, torch.Tensor) else d for d in data])
        ctx.tape, ctx.names, ctx.n_data = tape, names, n_data
        ctx.data = data
        ctx.params = params if need else None
        ctx.pshapes = [tuple(p.shape) for p in params]
        ctx.pdev = dev
        # only the identity of the output is kept (keeping the tensor would tie out.grad_fn -> ctx -> out into a cycle that
        # holds the whole tape of a forward that is never back-propagated until the cyclic GC runs)
        ctx.out_key, ctx.out_shape = S._Bwd.key(out), tuple(out.shape)
        return out.view(out.shape) if need else out          # the tape's closures hold `out` itself: hand autograd an alias

    @staticmethod
    def backward(ctx, dout):
        n_data, names = ctx.n_data, ctx.names
        need = ctx.needs_input_grad[4:]
        with dev_guard(ctx.pdev):
            sizes = [int(torch.Size(s).numel()) for s in ctx.pshapes]
            flat = torch.zeros(sum(sizes), device=ctx.pdev, dtype=torch.float32)
            pg, views, off = {}, [], 0
            for i, (name, shp, n) in enumerate(zip(names, ctx.pshapes, sizes)):
                v = flat[off:off + n].view(shp)
                off += n
                views.append(v)
                pg[name] = v if need[n_data + i] else None
            bw = S._Bwd(pg)
            dout = _prep(dout)
            bw.seed_key(ctx.out_key, dout if tuple(dout.shape) == ctx.out_shape else dout.view(ctx.out_shape))
            ctx.tape.backward(bw)
            if parallel.enabled():
                # the bucket may travel asynchronously only if autograd ADOPTS its views as .grad: parameters that already hold
                # a gradient (accumulation, zero_grad(set_to_none=False)) or a stack that already produced a bucket in this
                # backward pass (its views get summed in place) need the reduced values before autograd touches them
                params = ctx.params or ()
                first = not any(p.grad is not None for p in params) and parallel.first_bucket_of(params)
                parallel.bucket_ready(flat, blocking=not first)
            dgrads = []
            for i, d in enumerate(ctx.data):
                if isinstance(d, torch.Tensor) and need[i]:
                    g = bw.take(d.detach())
                    dgrads.append(g if g is not None else torch.zeros_like(d))
                else:
                    dgrads.append(None)
            pgrads = [v if need[n_data + i] else None for i, v in enumerate(views)]
        ctx.tape = None
        ctx.params = None
        return (None, None, None, None, *dgrads, *pgrads)


def drop_p_of(module: torch.nn.Module) -> float:
    """Dropout probability of a stack.  Modules built here record it; a module unpickled from a REFERENCE checkpoint
    (whole-module torch.save, cannon/test_photospectra.py:153) only has the reference's attributes, so it is recovered from
    the first Dropout / MultiheadAttention layer inside and cached."""
    p = module.__dict__.get("_drop_p")
    if p is None:
        p = 0.0
        for sub in module.modules():
            if isinstance(sub, torch.nn.Dropout):
                p = float(sub.p)
                break
            if isinstance(sub, torch.nn.MultiheadAttention):
                p = float(sub.dropout)
                break
        module.__dict__["_drop_p"] = p
    return float(p)


def model_dim_of(module: torch.nn.Module) -> int:
    """Width of a stack (recorded at construction, or read off the first LayerNorm of an unpickled reference module)."""
    d = module.__dict__.get("_model_dim")
    if d is None:
        d = next(int(sub.normalized_shape[0]) for sub in module.modules() if isinstance(sub, torch.nn.LayerNorm))
        module.__dict__["_model_dim"] = d
    return int(d)


def check_geometry(module: torch.nn.Module) -> None:
    """The fused kernels compute model_dim 32 with 4 heads of head_dim 8 (what every reference script builds).  The number of
    heads is NOT visible in any parameter shape (in_proj_weight is [3D, D] whatever it is), so it is checked on the
    nn.MultiheadAttention containers themselves — also for modules unpickled from a reference checkpoint — and anything else
    raises instead of silently computing with 4 heads (TransformerBlock ctor, util_layers.py:265-271)."""
    ok = module.__dict__.get("_geom_ok")
    if ok is None:
        ok = True
        for name, sub in module.named_modules():
            if isinstance(sub, torch.nn.MultiheadAttention) and (sub.embed_dim != 32 or sub.num_heads != 4):
                ok = (f"{name or type(module).__name__}: embed_dim={sub.embed_dim}, num_heads={sub.num_heads}")
                break
        module.__dict__["_geom_ok"] = ok
    if ok is not True:
        raise NotImplementedError(
            f"VAESNe-B200 kernels support model_dim=32 with num_heads=4 (head_dim 8), the geometry of every reference script; got {ok}")


def run_stack(module: torch.nn.Module, runner: Callable, data: Sequence):
    check_geometry(module)
    names, params = zip(*module.named_parameters())
    drop_p = drop_p_of(module) if module.training else 0.0
    return _StackFn.apply(runner, names, len(data), drop_p, *data, *params)


# ------------------------------------------------------------------------------------------------
class _LatentFn(torch.autograd.Function):
    """(bott_0..bott_{M-1}) -> z [M,K,B,T,Z], lat [M*K,B] (or an empty tensor), mu_0.., s_0..

    PhotometricVAE.py:53-54,162-163 / SpectraVAE.py:48-49,152-153 / losses.py:53-54."""

    @staticmethod
    def forward(ctx, fams, T, fam_prior, pz_mu, pz_s, want_lat, noises, *botts):
        M = len(botts)
        with dev_guard(botts[0].device):
            z, mus, ss, lat, pi = P.latent_fwd([b.detach() for b in botts], noises, fams, T, fam_prior, pz_mu, pz_s, want_lat)
        ctx.cfg = (fams, T, fam_prior, pz_mu, pz_s, want_lat, noises, pi)
        ctx.botts = [b.detach() for b in botts]
        if lat is None:
            lat = z.new_empty(0)
        else:
            lat = lat.view(M * noises[0].shape[0], -1)
        ctx.set_materialize_grads(False)
        return (z, lat, *mus, *ss)

    @staticmethod
    def backward(ctx, dz, dlat, *dms):
        fams, T, fam_prior, pz_mu, pz_s, want_lat, noises, pi = ctx.cfg
        M = len(ctx.botts)
        dmu = [_prep(g) for g in dms[:M]]
        ds = [_prep(g) for g in dms[M:]]
        with dev_guard(ctx.botts[0].device):
            dbotts = P.latent_bwd(ctx.botts, noises, fams, T, fam_prior, pz_mu, pz_s, _prep(dz),
                                  _prep(dlat) if (want_lat and dlat is not None) else None, pi,
                                  dmu if any(g is not None for g in dmu) else None,
                                  ds if any(g is not None for g in ds) else None)
        return (None, None, None, None, None, None, None, *dbotts)


def latent_step(botts, noises, fams, T, fam_prior=0, pz_mu=None, pz_s=None, want_lat=False):
    M = len(botts)
    out = _LatentFn.apply(list(fams), T, fam_prior, pz_mu, pz_s, want_lat, [_prep(n) for n in noises], *[_prep(b) for b in botts])
    z, lat = out[0], out[1]
    return z, (lat if want_lat else None), list(out[2:2 + M]), list(out[2 + M:2 + 2 * M])


# ------------------------------------------------------------------------------------------------
class LikSpec:
    """Likelihood of one modality: data, mask, family, fp32(1 + big), llik_scaling."""
    __slots__ = ("x", "mask", "fam", "scale_masked", "scaling")

    def __init__(self, x, mask, fam, scale_masked, scaling):
        self.x, self.mask, self.fam, self.scale_masked, self.scaling = _prep(x, torch.float32), _prep(mask), fam, scale_masked, scaling


class _IwaeFn(torch.autograd.Function):
    """objective = sum_b ( logsumexp_r (lat[r,b] + sum_d scaling_d * loglik_d[r,b]) - log R )   losses.py:55-62,93"""

    @staticmethod
    def forward(ctx, specs: List[LikSpec], lat, *locs):
        R, B = locs[0].shape[0], locs[0].shape[1]
        lpx = torch.empty(R, B, device=locs[0].device, dtype=torch.float32)
        locs = [_prep(l.detach()) for l in locs]
        for i, (sp, loc) in enumerate(zip(specs, locs)):
            P.loglik_fwd(loc, sp.x, sp.mask, sp.fam, sp.scale_masked, sp.scaling, lpx, accumulate=i > 0)
        lat_c = _prep(lat.detach()) if lat is not None and lat.numel() else None
        obj, w, _ = P.iwae_combine(lat_c, lpx)
        ctx.specs, ctx.locs, ctx.w, ctx.has_lat = specs, locs, w, lat_c is not None
        return obj

    @staticmethod
    def backward(ctx, g):
        g = _prep(g)
        dlat = P.scale(ctx.w, 1.0, g) if ctx.has_lat else None
        dlocs = [P.loglik_bwd(loc, sp.x, sp.mask, sp.fam, sp.scale_masked, sp.scaling, ctx.w, 1.0, g)
                 for sp, loc in zip(ctx.specs, ctx.locs)]
        return (None, dlat, *dlocs)


def iwae_objective(specs, lat, locs):
    return _IwaeFn.apply(specs, lat, *locs)


class _ElboFn(torch.autograd.Function):
    """objective = mean_{k,b} sum_l scaling*loglik - mean_b sum_{t,z} KL(q||p)   losses.py:16-24"""

    @staticmethod
    def forward(ctx, spec: LikSpec, fam_q, pz_mu, pz_s, loc, mu, s):
        K, B = loc.shape[0], loc.shape[1]
        loc, mu, s = _prep(loc.detach()), _prep(mu.detach()), _prep(s.detach())
        lpx = torch.empty(K, B, device=loc.device, dtype=torch.float32)
        P.loglik_fwd(loc, spec.x, spec.mask, spec.fam, spec.scale_masked, spec.scaling, lpx, False)
        kld = P.kl_fwd(mu, s, fam_q, pz_mu, pz_s)
        ctx.saved = (spec, fam_q, pz_mu, pz_s, loc, mu, s)
        ctx.parts = (lpx, kld)
        return P.elbo_combine(lpx, kld)

    @staticmethod
    def backward(ctx, g):
        spec, fam_q, pz_mu, pz_s, loc, mu, s = ctx.saved
        K, B = loc.shape[0], loc.shape[1]
        g = _prep(g)
        dloc = P.loglik_bwd(loc, spec.x, spec.mask, spec.fam, spec.scale_masked, spec.scaling, None, 1.0 / (K * B), g)
        dmu, ds = P.kl_bwd(mu, s, fam_q, pz_mu, pz_s, -1.0 / B, g)
        return (None, None, None, None, dloc, dmu, ds)


def elbo_objective(spec, fam_q, pz_mu, pz_s, loc, mu, s):
    return _ElboFn.apply(spec, fam_q, pz_mu, pz_s, loc, mu, s)


# ------------------------------------------------------------------------------------------------
class _InfoNCEFn(torch.autograd.Function):
    """-(CE(z1n z2n^T / tau, arange) + CE(z2n z1n^T / tau, arange)) / 2 with zXn = F.normalize(zX)   losses.py:98-110
    (single process: rows and columns are the same batch).  Every arithmetic step is a kernel of csrc/extra.cu."""

    @staticmethod
    def forward(ctx, z1, z2, temperature):
        z1, z2 = _prep(z1.detach(), torch.float32), _prep(z2.detach(), torch.float32)
        with dev_guard(z1.device):
            n = z1.shape[0]
            inv_tau = 1.0 / float(temperature)
            y1, i1 = P.l2norm_fwd(z1)
            y2, i2 = P.l2norm_fwd(z2)
            l1, lse1 = P.ce_rows_fwd(y1, y2, inv_tau)
            l2, lse2 = P.ce_rows_fwd(y2, y1, inv_tau)
            out = P.sum_scale(l1, l2, -0.5 / n)
        ctx.saved = (y1, i1, y2, i2, lse1, lse2, inv_tau, n)
        return out

    @staticmethod
    def backward(ctx, g):
        y1, i1, y2, i2, lse1, lse2, inv_tau, n = ctx.saved
        g = _prep(g)
        with dev_guard(y1.device):
            w = -0.5 / n
            dy1 = torch.empty_like(y1); dy2 = torch.empty_like(y2)
            P.ce_rows_bwd(y1, y2, inv_tau, 0, lse1, w, g, dA=dy1, dB=dy2)                               # rows = z1, columns = z2
            P.ce_rows_bwd(y2, y1, inv_tau, 0, lse2, w, g, dA=dy2, dA_acc=True, dB=dy1, dB_acc=True)     # rows = z2, columns = z1
            return P.l2norm_bwd(y1, i1, dy1), P.l2norm_bwd(y2, i2, dy2), None


def infonce_objective(z1, z2, temperature):
    return _InfoNCEFn.apply(z1, z2, temperature)


class _L2NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z):
        z = _prep(z.detach(), torch.float32)
        with dev_guard(z.device):
            y, inv = P.l2norm_fwd(z)
        ctx.saved = (y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved
        with dev_guard(y.device):
            return P.l2norm_bwd(y, inv, _prep(dy, torch.float32))


class _CeRowsFn(torch.autograd.Function):
    """sum_i [ logsumexp_j(A_i . B_j / tau) - A_i . B_{i + off} / tau ]: the data-parallel form (columns = the gathered batch)."""

    @staticmethod
    def forward(ctx, A, Bm, inv_tau, off):
        A, Bm = _prep(A.detach(), torch.float32), _prep(Bm.detach(), torch.float32)
        with dev_guard(A.device):
            loss, lse = P.ce_rows_fwd(A, Bm, inv_tau, off)
            out = P.sum_scale(loss, None, 1.0)
        ctx.saved = (A, Bm, lse, inv_tau, off)
        return out

    @staticmethod
    def backward(ctx, g):
        A, Bm, lse, inv_tau, off = ctx.saved
        with dev_guard(A.device):
            dA, dB = torch.empty_like(A), torch.empty_like(Bm)
            P.ce_rows_bwd(A, Bm, inv_tau, off, lse, 1.0, _prep(g), dA=dA, dB=dB)
        return dA, dB, None, None


def l2normalize(z):
    return _L2NormFn.apply(z)


def ce_rows_sum(A, Bm, inv_tau, off=0):
    return _CeRowsFn.apply(A, Bm, inv_tau, off)
