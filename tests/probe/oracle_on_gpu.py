"""Dev probe (test infrastructure, not a bench line): the reference ALGORITHM (oracle port, plain eager PyTorch ops, fp32)
run on the same B200 — the "stock PyTorch on a GPU" bar next to the CPU baseline.  B is small because eager attention
materialises P [rows*4, 982, 982] for every layer.   python tests/probe/oracle_on_gpu.py [B ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "vaesne-dev_b200"), ROOT]
import torch
import bench
from oracle import vaesne_oracle as O

dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = False
for B in [int(b) for b in (sys.argv[1:] or ["8", "16"])]:
    model = bench.build_model("cpu", 0.1)
    params = {k: v.detach().clone().to(dev).requires_grad_(v.is_floating_point() and "_pz_params" not in k) for k, v in model.state_dict().items()}
    cfg = O.MMVAEConfig([O.VAEConfig("photometry", 4, 4), O.VAEConfig("spectra", 4, 4)], beta=0.5)
    cfg.apply_scaling()
    opt = torch.optim.AdamW([p for p in params.values() if p.requires_grad], lr=1e-3)
    xs = [[tuple(t.to(dev) for t in mod) for mod in bench.synth_batch(B, 100 + i)] for i in range(2)]
    def step(i):
        x = xs[i % 2]
        us = [torch.empty(bench.KS, B, 4, 4, device=dev).uniform_(-0.999, 0.999) for _ in range(2)]
        for v in params.values():
            v.grad = None
        loss = -O.m_iwae(params, cfg, x, us, dropout=0.1)
        loss.backward()
        opt.step()
    try:
        for i in range(2):
            step(i)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 4
        for i in range(n):
            step(i)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
        print(f"oracle port on cuda, B={B}: {dt * 1e3:.1f} ms/step = {B / dt:.1f} samples/s; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    except torch.cuda.OutOfMemoryError as e:
        print(f"B={B}: out of memory", flush=True)
        break
