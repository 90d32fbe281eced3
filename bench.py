#!/usr/bin/env python
"""bench.py — train samples/s of the VAESNe mmVAE step (fwd + bwd + IW-ELBO + AdamW) on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    (the reference algorithm on the host CPU cores)

Workload (BASELINE.json configs[3], the one the metric is quoted on): the ZTF photometry+spectra
mixture-of-experts VAE of cannon/ZTF_photospect.py:76-119 — 2 bands, latent 4x4, model_dim 32, 4 heads,
4 layers, K=8 importance samples, beta=0.5, spectra-encoder context self-attention, dropout 0.1 (train
mode), AdamW lr 1e-3 — on synthetic Goldstein/ZTF-shaped batches (60 photometry points, 982 spectrum
bins), weak scaling with a fixed per-GPU batch.

One JSON line on stdout (rank 0).  `value` = samples/s with the batches resident in HBM; `e2e` = the same
step through the public API (VAESNe.training_util.training_step) from pinned host batches with the loss
read back every step.  `roofline` is for the dominant kernel (timed with CUDA events on the launching
stream in an extra instrumented pass), `cpu_baseline` is the oracle port timed on the host cores."""
from __future__ import annotations

import argparse
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
def synth_batch(B, seed, num_bands=2):
    from oracle import vaesne_oracle as O
    return [O.synth_photometry(B, LP, num_bands, seed=seed), O.synth_spectra(B, LS, seed=seed)]


def build_model(device, dropout=0.1):
    from VAESNe.PhotometricVAE import PhotometricVAE
    from VAESNe.SpectraVAE import SpectraVAE
    from VAESNe.mmVAE import photospecMMVAE
    torch.manual_seed(1)
    pv = PhotometricVAE(num_bands=2, latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=4,
                        dropout=dropout, selfattn=False, beta=0.5)
    sv = SpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=4,
                    dropout=dropout, selfattn=True, beta=0.5)
    return photospecMMVAE([pv, sv], beta=0.5).to(device)


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
    torch.set_num_threads(cores)
    t1 = time_cpu(1, 1, 1)                                   # calibrate: seconds per sample-step
    budget = 150.0
    B = int(max(1, min(16, budget / ((args.steps + args.warmup) * t1))))
    t = time_cpu(B, args.steps, args.warmup)
    v = B / t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(B), "per_step_batch": B, "K": KS},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{args.steps} steps of B={B} (oracle port of the reference, torch CPU fp32, dropout 0.1, AdamW)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(B):
    return (f"ZTF_photospect mmVAE train step (2 bands, K={KS}, beta=0.5, spectra-encoder selfattn, dropout 0.1, AdamW lr 1e-3), "
            f"Lp={LP}, Ls={LS}, per-GPU batch {B}")


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
        traffic_tab = {}
        try:
            traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        except Exception:
            pass
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
                                  "share_of_step": lv["total_ms"] / tot, "avg_ms": lv["avg_ms"],
                                  "note": "inputs partly L2-resident inside the step; tests/probe/lin_bench.py times the same kernels cold"}
        if rank == 0:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "kernel_breakdown.json"), "w") as f:
                json.dump({"|".join(map(str, k)): v for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["total_ms"])}, f, indent=1)

    # ---- encode latents/s (BASELINE.json's second metric; 1 GPU): VAE.encode on resident batches ------------
    enc = None
    if world == 1:
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

    if rank != 0:
        torch.distributed.destroy_process_group()
        return

    # ---- CPU baseline (oracle port on the host cores), bounded sample -----------------------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        t1 = time_cpu(1, 1, 1)
        Bc = int(max(1, min(8, 12.0 / t1)))
        t = time_cpu(Bc, 1, 0)
        cpu = {"value": Bc / t, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"1 step of B={Bc} after a B=1 warm-up (oracle port of the reference algorithm, torch CPU fp32, dropout 0.1, AdamW)"}

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
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "encode": enc}
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
