import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "vaesne-dev_b200"), ROOT]
import torch
from VAESNe import _ops as P
dev = "cuda"
def per_element(L, N):
    g = torch.Generator().manual_seed(L)
    qkv = torch.randn(N, L, 96, generator=g).to(dev); dO = torch.randn(N, L, 32, generator=g).to(dev); dqkv = torch.empty_like(qkv)
    q, k, v = qkv[..., :32], qkv[..., 32:64], qkv[..., 64:]
    def step():
        O, LSE = P.attn_fwd(q, k, v, None)
        P.attn_bwd(q, k, v, None, O, LSE, dO, dqkv[..., :32], dqkv[..., 32:64], dqkv[..., 64:])
    for _ in range(2): step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(5): step()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 5 / (N * 4.0 * L * L) * 1e6
print(os.environ.get("VAESNE_TC_MIN"), {L: round(per_element(L, max(256, int(4096 * (60.0 / L) ** 2))), 4) for L in (64, 96, 128, 160, 192, 224, 255, 256, 384)})
