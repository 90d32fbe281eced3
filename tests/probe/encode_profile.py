"""Dev probe: VAE.encode throughput (BASELINE's second metric) against the batch size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "vaesne-dev_b200"), ROOT]
import torch
import bench
dev = torch.device("cuda:0")
model = bench.build_model(dev)
for EB in (512, 2048, 8192):
    xd = [tuple(t.to(dev) for t in mod) for mod in bench.synth_batch(EB, 77)]
    for name, vae, x in (("photometry", model.vaes[0], xd[0]), ("spectra", model.vaes[1], xd[1])):
        for _ in range(3):
            vae.encode(x)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(5):
            vae.encode(x)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"batch {EB:5d} {name:10s}: {ms:8.3f} ms  {EB / ms * 1e3:10.0f} latents/s", flush=True)
