#!/usr/bin/env python
"""Times the UNMODIFIED reference (baseline/_ref/VAESNe, installed by baseline/install_ref.py) through its own public API:
`VAESNe.training_util.training_step(model, torch.optim.AdamW, batches, loss_fn, multimodal)` — train mode, dropout as shipped,
the per-step host->device copies and the loss read-back the reference does itself.  Runs in its OWN interpreter (bench.py
starts it as a child process) because the reference package and the product package share the name `VAESNe`.

    python baseline/ref_arm.py --config mmvae_ztf --device cpu --batch 4 --steps 3 --warmup 1

Prints one JSON object: samples/s (median step), per-step seconds, batch, device, threads, peak device memory."""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="mmvae_ztf")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--budget-s", type=float, default=0.0, help="if > 0: one B=1 calibration step, then the largest batch <= --batch "
                    "for which warmup+steps fit the budget")
    args = ap.parse_args()
    if not os.path.exists(os.path.join(REF, "VAESNe", "mmVAE.py")):
        print(json.dumps({"unavailable": "baseline/_ref is empty: run python baseline/install_ref.py where /root/reference exists"}))
        return
    # the reference package first; the product package must not be importable here
    sys.path[:] = [os.path.join(HERE, "stubs"), REF, ROOT] + [p for p in sys.path if "vaesne-dev_b200" not in p and p not in ("", ROOT)]
    import torch
    import VAESNe
    assert os.path.realpath(os.path.dirname(VAESNe.__file__)).startswith(os.path.realpath(REF)), VAESNe.__file__
    import bench_common as BC
    cores = os.cpu_count() or 1
    if args.device == "cpu":
        torch.set_num_threads(cores)
    dev = torch.device(args.device)
    ns = BC.Namespace()
    torch.manual_seed(1)
    model, loss_fn, make_batch, multimodal = BC.build_config(args.config, ns, args.dropout)
    model = model.to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=BC.CONFIGS[args.config]["lr"])

    def step(B, seed):
        x = make_batch(B, seed)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        ns.training_step(model, opt, [x], loss_fn=loss_fn, multimodal=multimodal)       # ends with loss.cpu().item(): synchronised
        return time.perf_counter() - t0

    B = args.batch
    calib = None
    if args.budget_s > 0:
        calib = step(1, 99)
        B = int(max(1, min(args.batch, args.budget_s / ((args.steps + args.warmup) * calib))))
    times = []
    for i in range(args.warmup + args.steps):
        dt = step(B, 100 + i)
        if i >= args.warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    mean = sum(times) / len(times)
    out = {"config": args.config, "device": args.device, "batch": B, "steps": args.steps, "warmup": args.warmup, "dropout": args.dropout,
           "s_per_step": mean, "s_per_step_median": med, "samples_per_s": B / mean, "threads": torch.get_num_threads(), "cores": cores,
           "calibration_s_per_sample": calib, "torch": torch.__version__,
           "peak_mem_gib": (round(torch.cuda.max_memory_allocated() / 2 ** 30, 2) if dev.type == "cuda" else None),
           "package": os.path.relpath(os.path.dirname(VAESNe.__file__), ROOT)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
