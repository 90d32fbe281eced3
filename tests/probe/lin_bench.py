"""Dev probe (test infrastructure): achieved HBM bandwidth of the linear kernels at the bench shape (1.0 M tokens)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "vaesne-dev_b200"), ROOT]
import torch
from VAESNe import _ops as P

dev = "cuda"
T = int(os.environ.get("T", 1005568))
g = torch.Generator().manual_seed(0)
def rnd(*s): return torch.randn(*s, generator=g).to(dev)
X, R, dY = rnd(T, 32), rnd(T, 32), rnd(T, 32)
W32, b32 = rnd(32, 32) / 6, rnd(32)
W96, b96 = rnd(96, 32) / 6, rnd(96)
gam, bet = rnd(32), rnd(32)
S = torch.empty(T, 32, device=dev); H = torch.empty(T, 32, device=dev)
dY96 = rnd(T, 96)
seed = torch.tensor([1234], dtype=torch.int64, device=dev)
drop = P.Drop(0.1, seed, 5)
dW32 = torch.zeros(32, 32, device=dev); db32 = torch.zeros(32, device=dev); dW96 = torch.zeros(96, 32, device=dev); db96 = torch.zeros(96, device=dev)
dg = torch.zeros(32, device=dev); dbe = torch.zeros(32, device=dev)
dX = torch.empty(T, 32, device=dev); dR = torch.empty(T, 32, device=dev)
Y = torch.empty(T, 32, device=dev); Y96 = torch.empty(T, 96, device=dev)

cases = [
    ("fwd_ln   (X,R -> S,Y)        512 B/tok", 512, lambda: P.lin_fwd(X, W32, b32, R=R, gamma=gam, beta=bet, S=S, drop=drop, Y=Y)),
    ("fwd 32   (X -> Y)            256 B/tok", 256, lambda: P.lin_fwd(X, W32, b32, Y=Y)),
    ("fwd gelu (X -> H,Y)          384 B/tok", 384, lambda: P.lin_fwd(X, W32, b32, act=P.ACT_GELU, H=H, Y=Y)),
    ("fwd 96   (X -> Y96)          512 B/tok", 512, lambda: P.lin_fwd(X, W96, b96, Y=Y96)),
    ("bwd_ln   (dY,S,X -> dR,dX)   640 B/tok", 640, lambda: P.lin_bwd(dY, X, W32, S=S, gamma=gam, dgamma=dg, dbeta=dbe, dR=dR, drop=drop, dW=dW32, db=db32, dX=dX)),
    ("bwd 32   (dY,X -> dX)        384 B/tok", 384, lambda: P.lin_bwd(dY, X, W32, dW=dW32, db=db32, dX=dX)),
    ("bwd gelu (dY,H,X -> dX)      512 B/tok", 512, lambda: P.lin_bwd(dY, X, W32, act=P.ACT_GELU, A=H, dW=dW32, db=db32, dX=dX)),
    ("bwd 96   (dY96,X -> dX)      640 B/tok", 640, lambda: P.lin_bwd(dY96, X, W96, dW=dW96, db=db96, dX=dX)),
    ("bwd 32 dX only (dY -> dX)    256 B/tok", 256, lambda: P.lin_bwd(dY, None, W32, dX=dX)),
    ("bwd 32 dW only (dY,X -> )    256 B/tok", 256, lambda: P.lin_bwd(dY, X, W32, dW=dW32, db=db32)),
]
for name, bpt, fn in cases:
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"{name}: {ms:.3f} ms  {T * bpt / ms / 1e6:7.0f} GB/s  ({T * bpt / ms / 1e6 / 6544 * 100:.0f}% of measured HBM peak)", flush=True)
