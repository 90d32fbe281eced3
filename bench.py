#!/usr/bin/env python
"""bench.py — train samples/s of the VAESNe mmVAE step (fwd + bwd + IW-ELBO + AdamW) on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    (the UNMODIFIED reference on the host CPU cores)

Workload (BASELINE.json configs[3], the one the metric is quoted on): the ZTF photometry+spectra
mixture-of-experts VAE of cannon/ZTF_photospect.py:76-119 — 2 bands, latent 4x4, model_dim 32, 4 heads,
4 layers, K=8 importance samples, beta=0.5, spectra-encoder context self-attention, dropout 0.1 (train
mode), AdamW lr 1e-3 — on synthetic Goldstein/ZTF-shaped batches (60 photometry points, 982 spectrum
bins), weak scaling with a fixed per-GPU batch.

One JSON line on stdout (rank 0).  `value` = samples/s with the batches resident in HBM; `e2e` = the same
step through the public API (VAESNe.training_util.training_step) from pinned host batches with the loss
read back every step.  `roofline` is for the dominant kernel (timed with CUDA events on the launching
stream in an extra instrumented pass).  Next to it, at N=1: `cpu_baseline` (the unmodified reference on the
host cores), `eager_gpu_baseline` (the same unmodified reference as stock eager PyTorch on this GPU),
`configs` (samples/s of BASELINE.json's other configurations), `batch_sweep` (per-GPU batch 16..512) and the
encode / reconstruct rates; at N>1: `dp_check` (all-reduced buckets == sum over ranks of the local gradients).

The reference arm runs baseline/_ref (pip-installed by baseline/install_ref.py from /root/reference, git-ignored,
shipped to the GPU box) in a child interpreter through baseline/ref_arm.py; only if that copy is missing does it
fall back to the oracle port (`cpu_baseline.kind` says which).  The product arm never imports oracle/."""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "vaesne-dev_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

import bench_common as BC  # noqa: E402

LP, LS = 60, 982
KS = 8               # importance samples (ZTF_photospect.py:117)
METRIC = "train samples/s (mmVAE fwd+bwd+IW-ELBO+AdamW)"
UNIT = "samples/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# per-sample algorithmic matmul FLOPs, fwd+bwd (3 x fwd; per row of one block
# 128*Lq^2 + 128*Lq*Lc + 16384*Lq + 4096*Lc)
def block_flops(Lq, Lc):
    return 128 * Lq * Lq + 128 * Lq * Lc + 16384 * Lq + 4096 * Lc


def step_flops_per_sample(K):
    enc_p = 4 * block_flops(8, LP)
    enc_s = 4 * (block_flops(8, LS + 1) + 128 * (LS + 1) ** 2 + 8192 * (LS + 1))     # + context self-attention
    dec_p = 4 * block_flops(LP, 4)
    dec_s = 4 * block_flops(LS, 5)
    return 3 * (enc_p + enc_s + 2 * K * (dec_p + dec_s))


# ------------------------------------------------------------------------------------------------
synth_batch = BC.synth_batch


def build_model(device, dropout=0.1):
    torch.manual_seed(1)
    model, _, _, _ = BC.build_config("mmvae_ztf", BC.Namespace(), dropout)
    return model.to(device)


def run_ref_arm(**kw):
    """baseline/ref_arm.py in a child interpreter (the reference package shares the product package's name) -> dict or None."""
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "ref_arm.py")]
    for k, v in kw.items():
        cmd += ["--" + k.replace("_", "-"), str(v)]
    env = {k: v for k, v in os.environ.items() if k not in ("PYTHONPATH", "RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not line:
            return {"unavailable": (r.stderr or r.stdout)[-300:].replace("\n", " | ")}
        return json.loads(line[-1])
    except Exception as e:      # noqa: BLE001
        return {"unavailable": repr(e)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(s for s in sm if s > 0.5 * max(mx or [1]))  or sorted(sm)
        med = busy[len(busy) // 2] if busy else None
        return dict(sm_mhz=med, sm_max_mhz=(max(mx) if mx else None), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------
def cpu_reference_step(params, cfg, x, K, dropout):
    """One train step of the reference algorithm (oracle port) on the CPU: fwd + bwd + AdamW."""
    from oracle import vaesne_oracle as O
    us = [O.draw_noise("laplace", (K, x[0][0].shape[0], 4, 4)) for _ in range(2)]
    for v in params.values():
        v.grad = None
    loss = -O.m_iwae(params, cfg, x, us, dropout=dropout)
    loss.backward()
    return loss.detach()


def make_cpu_reference(dropout=0.1):
    from oracle import vaesne_oracle as O
    model = build_model("cpu", dropout)
    params = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "_pz_params" not in k) for k, v in model.state_dict().items()}
    cfg = O.MMVAEConfig([O.VAEConfig("photometry", 4, 4), O.VAEConfig("spectra", 4, 4)], beta=0.5)
    cfg.apply_scaling()
    opt = torch.optim.AdamW([p for p in params.values() if p.requires_grad], lr=1e-3)
    return params, cfg, opt


def time_cpu(B, steps, warmup, dropout=0.1):
    torch.set_num_threads(os.cpu_count() or 1)
    params, cfg, opt = make_cpu_reference(dropout)
    times = []
    for i in range(warmup + steps):
        x = synth_batch(B, 100 + i)
        t0 = time.perf_counter()
        cpu_reference_step(params, cfg, x, KS, dropout)
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:            # under torchrun only rank 0 runs the CPU arm; the other ranks exit without work
        return
    cores = os.cpu_count() or 1
    ref = run_ref_arm(config="mmvae_ztf", device="cpu", batch=16, steps=args.steps, warmup=args.warmup, budget_s=150)
    if ref and "samples_per_s" in ref:
        B, t, v, kind = ref["batch"], ref["s_per_step"], ref["samples_per_s"], "reference"
        sample = (f"{args.steps} steps of B={B} after {args.warmup} warm-up (unmodified reference package {ref['package']} through its own "
                  f"training_step + torch.optim.AdamW, torch {ref['torch']} CPU fp32, {ref['threads']} threads, dropout 0.1)")
    else:                    # baseline/_ref missing on this box: the oracle port of the same algorithm
        torch.set_num_threads(cores)
        t1 = time_cpu(1, 1, 1)
        B = int(max(1, min(16, 150.0 / ((args.steps + args.warmup) * t1))))
        t = time_cpu(B, args.steps, args.warmup)
        v, kind = B / t, "port"
        sample = f"{args.steps} steps of B={B} (oracle port of the reference, torch CPU fp32, dropout 0.1, AdamW); reference copy unavailable: {ref}"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(B), "per_step_batch": B, "K": KS,
                       "note": "same model / objective / optimiser as the product arm; the step is a bounded sample (smaller batch) of that workload"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(B):
    return (f"ZTF_photospect mmVAE train step (2 bands, K={KS}, beta=0.5, spectra-encoder selfattn, dropout 0.1, AdamW lr 1e-3), "
            f"Lp={LP}, Ls={LS}, per-GPU batch {B}")


# ------------------------------------------------------------------------------------------------
def _time_train(model, opt, loss_fn, batches, multimodal, steps, warmup, graph=False):
    """samples/s of training_step over device-resident batches (CUDA events on the launching stream)."""
    from VAESNe.training_util import training_step
    nb = len(batches)
    B = (batches[0][0][0] if multimodal else batches[0][0]).shape[0]
    training_step(model, opt, [batches[i % nb] for i in range(warmup)], loss_fn, multimodal=multimodal, cuda_graph=graph)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    training_step(model, opt, [batches[i % nb] for i in range(steps)], loss_fn, multimodal=multimodal, cuda_graph=graph)
    b.record(); torch.cuda.synchronize()
    return B * steps / (a.elapsed_time(b) * 1e-3)


def _to_dev(x, dev, multimodal):
    return [tuple(t.to(dev) for t in m) for m in x] if multimodal else tuple(t.to(dev) for t in x)


def run_configs(dev):
    """BASELINE.json configs 1, 2, 3, 5 through the product's public API: per-script batch (eager and as a replayed
    CUDA graph — those batches are launch-bound) and a machine-filling batch."""
    from VAESNe.optim import FusedAdamW
    ns = BC.Namespace()
    out = {}
    big = {"photometry_elbo": 16384, "spectra_elbo": 1024, "mmvae_goldstein": 512, "contrastive": 2048, "photo_end2end": 16384}
    for name in ("photometry_elbo", "spectra_elbo", "mmvae_goldstein", "contrastive", "photo_end2end"):
        c = BC.CONFIGS[name]
        torch.manual_seed(1)
        model, loss_fn, make, mm = BC.build_config(name, ns, 0.1)
        model = model.to(dev)
        opt = FusedAdamW(model.parameters(), lr=c["lr"], grad_average=False)
        rec = {"script": c["script"], "K": c["K"]}
        for tag, B, steps, graph in (("script_batch", c["batch"], 20, False), ("script_batch_cuda_graph", c["batch"], 20, True),
                                     ("large_batch", big[name], 5, False)):
            batches = [_to_dev(make(B, 300 + i), dev, mm) for i in range(2)]
            try:
                rec[tag] = {"batch": B, "samples_per_s": _time_train(model, opt, loss_fn, batches, mm, steps, 4 if graph else 3, graph)}
            except Exception as e:      # noqa: BLE001
                rec[tag] = {"batch": B, "error": repr(e)[:200]}
            del batches
        out[name] = rec
        del model, opt
        torch.cuda.empty_cache()
    return out


def run_batch_sweep(dev, headline):
    """mmVAE ZTF step at per-GPU batch 16 / 64 / 256 (+ the headline batch): where the step stops being launch-bound."""
    from VAESNe.losses import m_iwae
    from VAESNe.optim import FusedAdamW
    model = build_model(dev)
    opt = FusedAdamW(model.parameters(), lr=1e-3)
    loss_fn = lambda m, x: m_iwae(m, x, K=KS)      # noqa: E731
    out = []
    for B in (16, 64, 256):
        batches = [_to_dev(synth_batch(B, 500 + i), dev, True) for i in range(2)]
        rec = {"per_gpu_batch": B, "samples_per_s": _time_train(model, opt, loss_fn, batches, True, 8 if B < 256 else 5, 3)}
        if B <= 64:
            rec["samples_per_s_cuda_graph"] = _time_train(model, opt, loss_fn, batches, True, 8, 4, graph=True)
        out.append(rec)
        del batches
    out.append({"per_gpu_batch": headline[0], "samples_per_s": headline[1]})
    del model, opt
    torch.cuda.empty_cache()
    return out


def run_dp_small_batch(model, opt, loss_fn, dev, rank, world, timed):
    """SURVEY 8d's per-GPU batch sweep under data parallelism: 16 (cannon/ZTF_photospect.py:77, launch-bound; eager and replayed
    as one CUDA graph with the NCCL bucket all-reduces captured inside), 256 and 1024; max over ranks like the headline."""
    from VAESNe.training_util import training_step
    out = []
    for B, steps in ((16, 20), (256, 5), (1024, 3)):
        batches = [_to_dev(synth_batch(B, 7000 + 10 * rank + i), dev, True) for i in range(2)]
        rec = {"per_gpu_batch": B, "global_batch": B * world}
        for tag, graph in (("samples_per_s", False),) + ((("samples_per_s_cuda_graph", True),) if B == 16 else ()):
            def run(n, graph=graph):
                training_step(model, opt, [batches[i % 2] for i in range(n)], loss_fn, multimodal=True, cuda_graph=graph)
            try:
                run(4 if B == 16 else 2)
                ms = timed(run, steps)
                rec[tag] = world * B * steps / (ms * 1e-3)
            except Exception as e:      # noqa: BLE001
                rec[tag] = None
                rec[tag + "_error"] = repr(e)[:200]
        out.append(rec)
        del batches
        torch.cuda.empty_cache()
    return out


def run_dp_check(model, loss_fn, x, dev, world):
    """One backward with the data-parallel plumbing off (local gradients) and one with it on (per-stack buckets all-reduced
    asynchronously over NCCL), same noise and dropout seeds; the reduced gradients must equal the sum over ranks of the local
    ones (gathered and added in rank order in fp64).  Reported, not asserted: rel = max |a - b| / max |b| over all parameters."""
    from VAESNe import _ops as P, parallel
    import torch.distributed as dist

    def grads(enabled):
        import itertools
        from VAESNe import _stacks
        parallel.enable(enabled)
        P._SEED_CELLS.clear()                      # the dropout seed cell is re-drawn from torch's (re-seeded) generator
        _stacks._call_counter = itertools.count(1)  # ... and the dropout stream ids restart, so both passes draw the same masks
        torch.manual_seed(4242 + dist.get_rank())
        model.zero_grad(set_to_none=True)
        (-loss_fn(model, x)).backward()
        parallel.wait_all()
        torch.cuda.synchronize()
        return torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]).clone()

    local = grads(False)
    reduced = grads(True)
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local)
    want = torch.stack([t.double() for t in parts]).sum(0)
    err = (reduced.double() - want).abs().max().item()
    ref = want.abs().max().item()
    bitwise = bool(torch.equal(reduced, want.float()))
    model.zero_grad(set_to_none=True)
    return {"rel_err": err / max(ref, 1e-30), "max_abs_grad": ref, "n_params": int(local.numel()), "equals_fp64_sum_rounded": bitwise,
            "what": "all-reduced per-stack buckets vs sum over ranks of the local gradients (same seeds), NCCL"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=512, help="per-GPU batch (1 B200: 1973 samples/s @256, 2046 @512; saved activations ~0.14 GB per sample)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip configs / batch_sweep / eager_gpu_baseline / encode")
    ap.add_argument("--cuda-graph", type=int, default=0, help="1: replay the captured step (training_step(cuda_graph=True))")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    from VAESNe import _native, _ops as P, parallel
    from VAESNe.losses import m_iwae
    from VAESNe.optim import FusedAdamW
    from VAESNe.training_util import training_step

    rank, world, local = parallel.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    assert not _native.is_emulated()
    B = args.batch
    model = build_model(dev)
    opt = FusedAdamW(model.parameters(), lr=1e-3)
    loss_fn = lambda m, x: m_iwae(m, x, K=KS)
    nb = 4                                     # distinct resident batches, cycled
    host = [synth_batch(B, 1000 * rank + i) for i in range(nb)]
    pinned = [[tuple(t.pin_memory() for t in mod) for mod in b] for b in host]
    resident = [[tuple(t.to(dev) for t in mod) for mod in b] for b in host]
    h2d = sum(t.numel() * t.element_size() for mod in host[0] for t in mod)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(steps)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return t.item()

    # ---- device-resident throughput ---------------------------------------------------------
    mode = {"graph": bool(args.cuda_graph)}

    def graph_launches():
        return sum(g.replayed_launches for g in opt.__dict__.get("_vaesne_graphed", {}).values())

    def run_resident(steps):
        training_step(model, opt, [resident[i % nb] for i in range(steps)], loss_fn, multimodal=True, cuda_graph=mode["graph"])

    run_resident(args.warmup)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = _native.launch_count() + graph_launches()
    ms = timed(run_resident, args.steps)
    launches = _native.launch_count() + graph_launches() - l0
    clk = clocks.stop() if rank == 0 else None
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through the public API: pinned host batch in, loss out, every step -----------
    last = {}

    def run_e2e(steps):
        for i in range(steps):
            last["loss"] = training_step(model, opt, [pinned[i % nb]], loss_fn, multimodal=True, cuda_graph=mode["graph"])

    run_e2e(1)
    ms_e2e = timed(run_e2e, args.steps)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)

    # ---- instrumented pass: per-kernel CUDA-event timing for the roofline line ---------------------
    roof = None
    if not args.no_profile:
        P.PROFILER = P.Profiler()
        saved, mode["graph"] = mode["graph"], False          # per-kernel events need the eager path
        run_resident(2)
        mode["graph"] = saved
        summ = P.PROFILER.summary()
        P.PROFILER = None
        tot = sum(v["total_ms"] for v in summ.values())
        key, top = max(summ.items(), key=lambda kv: kv[1]["total_ms"])
        pk = peaks()
        name = key[0]
        # measured DRAM traffic per batch row (ncu --set full of the SAME build: the table carries the hash of the kernel sources and a
        # mismatched table is refused, so the figure cannot go stale silently; profiles/make_traffic.py regenerates it)
        traffic_tab, traffic_note = {}, None
        try:
            tab = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            import importlib.util
            spec = importlib.util.spec_from_file_location("vaesne_b200_build", os.path.join(PKG, "build.py"))
            bmod = importlib.util.module_from_spec(spec); spec.loader.exec_module(bmod)
            sha = bmod.source_hash()
            if tab.get("src_sha256") == sha:
                traffic_tab = tab
            else:
                traffic_note = f"profiles/r2_traffic.json was measured on kernel sources {str(tab.get('src_sha256'))[:12]}, these are {sha[:12]}: refused"
        except Exception as e:      # noqa: BLE001
            traffic_note = f"no traffic table: {e!r}"
        if name.startswith("attn"):
            Nb, Lq, Lk = key[1:]
            # algorithmic work per launch (the two attention matmuls, 2*8 FLOP each per score element;
            # backward = 2.5x forward): what `achieved` is computed from
            elems = float(Nb) * 4 * Lq * Lk
            flops = 4.0 * elems * 8 * (1.0 if name == "attn_fwd" else 2.5)
            ach = flops / (top["avg_ms"] * 1e-3) / 1e12
            # the unit that actually binds at head_dim 8: one MUFU.EX2 per score element per pass,
            # 16 per clock per SM (tests/probe/tc_rates.cu), at the SM clock seen under load
            mhz = (clk or {}).get("sm_mhz") or 1965.0
            ex2_rate = 148 * 16 * mhz * 1e6
            ex2_floor_ms = elems / ex2_rate * 1e3            # the fused backward also exponentiates each element once
            roof = {"kernel": f"{name}[N={Nb},Lq={Lq},Lk={Lk}] (attn_tc_* kernels: tcgen05 kind::tf32 + kind::f16)", "bound": "tensor",
                    "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                    "frac": ach / pk["tf_sust"],
                    "traffic": (lambda t: None if t is None else int(t * Nb))(traffic_tab.get("per_row", {}).get(f"{name}|{Lq}|{Lk}")),
                    "traffic_note": traffic_note,
                    "peak_source": pk["src"] + " bf16 sustained (kernel timed inside the step)",
                    "share_of_step": top["total_ms"] / tot, "avg_ms": top["avg_ms"],
                    "binding_unit": {"name": "MUFU.EX2 (32 tensor FLOP per exponential at head_dim 8)", "floor_ms": ex2_floor_ms,
                                     "frac_of_floor": ex2_floor_ms / top["avg_ms"]}}
        else:
            T, Kd, Nd = key[1:]
            mult = 1.0 if "fwd" in name else 2.0
            flops = 2.0 * T * Kd * Nd * mult
            ach = flops / (top["avg_ms"] * 1e-3) / 1e12
            roof = {"kernel": f"{name}[T={T},K={Kd},N={Nd}]", "bound": "tensor", "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                    "frac": ach / pk["tf_sust"], "traffic": None, "peak_source": pk["src"] + " bf16 sustained",
                    "share_of_step": top["total_ms"] / tot, "avg_ms": top["avg_ms"]}
        # the fused bandwidth-bound kernel with the largest share: algorithmic bytes (128 B per token per tensor touched)
        lin = [(k, v) for k, v in summ.items() if k[0] in ("lin_fwd_ln", "lin_bwd_ln")]     # well-defined traffic: 4 / 5 tensors
        if lin and roof is not None:
            lk, lv = max(lin, key=lambda kv: kv[1]["total_ms"])
            T, Kd, Nd = lk[1:]
            tensors = {"lin_fwd_ln": 4, "lin_bwd_ln": 5, "lin_fwd": 1 + Nd / 32.0, "lin_bwd": 2 + Nd / 32.0}[lk[0]]
            gbs = T * 128.0 * tensors / (lv["avg_ms"] * 1e-3) / 1e9
            roof["hbm_kernel"] = {"kernel": f"{lk[0]}[T={T},K={Kd},N={Nd}] (lin_tc_* kernels)", "bound": "hbm", "achieved": gbs, "peak": pk["hbm"],
                                  "unit": "GB/s", "frac": gbs / pk["hbm"], "bytes_per_token": 128.0 * tensors,
                                  "traffic": (lambda t: None if t is None else int(t * T))(traffic_tab.get("per_token", {}).get(lk[0])),
                                  "share_of_step": lv["total_ms"] / tot, "avg_ms": lv["avg_ms"],
                                  "note": "inputs partly L2-resident inside the step; tests/probe/lin_bench.py times the same kernels cold"}
        if rank == 0:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "kernel_breakdown.json"), "w") as f:
                json.dump({"|".join(map(str, k)): v for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["total_ms"])}, f, indent=1)

    # ---- encode latents/s (BASELINE.json's second metric; 1 GPU): VAE.encode on resident batches ------------
    enc = None
    if world == 1 and not args.no_extras:
        enc = {}
        # the encoders are launch-bound below a few thousand samples (photometry: ~60 kernels of ~20 us at batch 512), so the
        # headline figure uses batch 8192; batch 512 is reported next to it
        for EB, suffix in ((8192, ""), (512, "_b512")):
            xb = synth_batch(EB, 77)
            xd = [tuple(t.to(dev) for t in mod) for mod in xb]
            for name, vae, x in (("photometry", model.vaes[0], xd[0]), ("spectra", model.vaes[1], xd[1])):
                was_training = vae.training
                for _ in range(3):
                    vae.encode(x)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); a.record()
                for _ in range(5):
                    vae.encode(x)
                b.record(); torch.cuda.synchronize()
                enc[name + "_latents_per_s" + suffix] = EB * 5 / (a.elapsed_time(b) * 1e-3)
                if EB == 512:               # launch-bound there: the same call replayed as one CUDA graph (vae.graph_encode)
                    vae.graph_encode = True
                    for _ in range(4):
                        vae.encode(x)
                    torch.cuda.synchronize(); a.record()
                    for _ in range(20):
                        vae.encode(x)
                    b.record(); torch.cuda.synchronize()
                    enc[name + "_latents_per_s" + suffix + "_cuda_graph"] = EB * 20 / (a.elapsed_time(b) * 1e-3)
                    vae.graph_encode = False
                vae.train(was_training)
        enc["batch"] = 8192
        # paper-scale inference, reconstruct(data, K=100) — all four cross-modal decodes of K samples per object
        # (decoder layer-0 self-attention is shared across the K samples and both sources: the queries are the data's embeddings)
        RB, RK = 8, 100
        xr = [tuple(t[:RB].contiguous() for t in mod) for mod in xd]
        for _ in range(2):
            model.reconstruct(xr, K=RK)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(3):
            model.reconstruct(xr, K=RK)
        b.record(); torch.cuda.synchronize()
        enc["reconstruct_K100_objects_per_s"] = RB * 3 / (a.elapsed_time(b) * 1e-3)
        enc["reconstruct_batch"] = RB
        model.train()

    # ---- N > 1: the all-reduced gradient buckets equal the sum over ranks of the local gradients (hardware NCCL path) -----
    dp_check = None
    if world > 1:
        dp_check = run_dp_check(model, loss_fn, resident[0], dev, world)

    # ---- N > 1: the reference scripts' own batch (16 per GPU) is launch-bound; eager vs the step replayed as ONE CUDA graph
    #      with the NCCL bucket all-reduces captured inside it
    dp_small = None
    if world > 1 and not args.no_extras:
        dp_small = run_dp_small_batch(model, opt, loss_fn, dev, rank, world, timed)

    # ---- BASELINE.json's other configurations and the per-GPU batch sweep (1 GPU) --------------------------------------
    cfgs = sweep = None
    if world == 1 and not args.no_extras:
        del resident, pinned
        torch.cuda.empty_cache()
        cfgs = run_configs(dev)
        sweep = run_batch_sweep(dev, headline=(B, value))

    if rank != 0:
        _finish(opt, world)
        return

    # ---- the unmodified reference: stock eager PyTorch on this GPU, and on the host cores (bounded samples) --------------
    cpu = eager = None
    if world == 1 and not args.no_extras:
        del model, opt
        torch.cuda.empty_cache()
        r = run_ref_arm(config="mmvae_ztf", device="cuda", batch=16, steps=3, warmup=2)
        eager = ({"value": r["samples_per_s"], "unit": UNIT, "batch": r["batch"], "ms_per_step": r["s_per_step"] * 1e3,
                  "peak_mem_gib": r["peak_mem_gib"], "kind": "reference",
                  "what": f"unmodified reference ({r['package']}) on cuda: its own training_step + torch.optim.AdamW, fp32 eager, dropout 0.1; "
                          "batch 16 is the script's own (it materialises every 982x982 probability tensor: ~2.6 GB per sample)"}
                 if r and "samples_per_s" in r else {"unavailable": r})
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        r = run_ref_arm(config="mmvae_ztf", device="cpu", batch=8, steps=2, warmup=1, budget_s=30)
        if r and "samples_per_s" in r:
            cpu = {"value": r["samples_per_s"], "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"{r['steps']} steps of B={r['batch']} after a warm-up step (unmodified reference {r['package']}, its own training_step + "
                             f"torch.optim.AdamW, torch CPU fp32, {r['threads']} threads, dropout 0.1)"}
        else:
            t1 = time_cpu(1, 1, 1)
            Bc = int(max(1, min(8, 12.0 / t1)))
            t = time_cpu(Bc, 1, 0)
            cpu = {"value": Bc / t, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"1 step of B={Bc} after a B=1 warm-up (oracle port; reference copy unavailable: {r})"}

    fl = step_flops_per_sample(KS)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(B), "global_batch": B * world, "K": KS, "parallelism": f"dp{world}", "cuda_graph": mode["graph"],
                       "l2": "per-step working set (saved activations, GBs) exceeds the 126 MB L2; 4 distinct batches cycled; no explicit flush",
                       "algorithmic_gflop_per_sample": fl / 1e9, "peak_hbm_gib": round(torch.cuda.max_memory_allocated() / 2**30, 1)},
            "achieved_tflops_step": value * fl / 1e12,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps,
                    "last_loss": last.get("loss")},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "eager_gpu_baseline": eager,
            "configs": cfgs, "batch_sweep": sweep, "dp_check": dp_check, "dp_small_batch": dp_small, "encode": enc}
    print(json.dumps(line), flush=True)
    _finish(None, world)


def _finish(opt, world):
    """Leave without tearing NCCL down: destroy_process_group() after a CUDA graph has captured NCCL launches blocks in the
    watchdog (seen on the 2-GPU box: 'CudaEventDestroy hang'); everything is flushed, so the process simply exits."""
    if world > 1:
        sys.stdout.flush(); sys.stderr.flush()
        torch.cuda.synchronize()
        try:
            torch.distributed.barrier()
        except Exception:      # noqa: BLE001
            pass
        os._exit(0)


if __name__ == "__main__":
    main()
