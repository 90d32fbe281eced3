"""Dev probe (test infrastructure): pipelined linear backward with accumulating dR / dX at several token counts."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "vaesne-dev_b200"), ROOT]
import torch
from VAESNe import _ops as P

dev = "cuda"
for T in [int(x) for x in os.environ.get("TS", "300,4096,40000,262144,1005568").split(",")]:
    g = torch.Generator().manual_seed(T)
    def rnd(*s): return torch.randn(*s, generator=g).to(dev)
    X, R, dY = rnd(T, 32), rnd(T, 32), rnd(T, 32)
    W, b, gam, bet = rnd(32, 32) / 6, rnd(32), rnd(32), rnd(32)
    S = torch.empty(T, 32, device=dev); Y = torch.empty(T, 32, device=dev)
    seed = torch.tensor([99], dtype=torch.int64, device=dev)
    drop = P.Drop(0.1, seed, 3)
    P.lin_fwd(X, W, b, R=R, gamma=gam, beta=bet, S=S, drop=drop, Y=Y)
    if os.environ.get('MARK'): S[:, 0] = torch.arange(T, device=dev, dtype=torch.float32)
    outs = {}
    for racc, xacc in [(False, False), (True, False), (False, True), (True, True)]:
        dW = torch.zeros(32, 32, device=dev); db = torch.zeros(32, device=dev); dg = torch.zeros(32, device=dev); dbe = torch.zeros(32, device=dev)
        dR = torch.full((T, 32), 0.25 if racc else 7.0, device=dev); dX = torch.full((T, 32), 0.5 if xacc else -3.0, device=dev)
        P.lin_bwd(dY, X, W, S=S, gamma=gam, dgamma=dg, dbeta=dbe, dR=dR, dR_acc=racc, drop=drop, dW=dW, db=db, dX=dX, dX_acc=xacc)
        torch.cuda.synchronize(); print('   ok', T, racc, xacc, flush=True)
        outs[(racc, xacc)] = (dR - (0.25 if racc else 0.0), dX - (0.5 if xacc else 0.0), dW, db, dg, dbe)
    ref = outs[(False, False)]
    for key, o in outs.items():
        errs = [float((a - r).abs().max() / (r.abs().max() + 1e-30)) for a, r in zip(o, ref)]
        print(f"T={T} dR_acc={key[0]} dX_acc={key[1]}: max rel diff vs overwrite mode {max(errs):.2e}  [dR dX dW db dg dbe] = " + " ".join(f"{e:.1e}" for e in errs), flush=True)
        if max(errs) > 1e-3:
            for name, a_, r_ in zip(["dR", "dX"], o[:2], ref[:2]):
                bad = ((a_ - r_).abs().amax(dim=1) > 1e-3 * r_.abs().max()).nonzero().flatten()
                if bad.numel():
                    tiles = torch.unique(bad // 128)
                    if name == "dR":
                        t0 = int(tiles[0]); rows = bad[(bad // 128) == t0]
                        print("      warps of the bad rows:", sorted(set(((rows % 128) // 32).tolist())))
                        for d in (-3, -2, -1, 1, 2, 3, 4):
                            t1 = t0 + d * 148
                            if 0 <= t1 < T // 128:
                                other = r_[t1 * 128 + (rows % 128)]
                                err = float((a_[rows] - other).abs().max() / (other.abs().max() + 1e-30))
                                print(f"      vs reference rows of tile {t1} ({d:+d} CTA steps): {err:.2e}")
                    print(f"    {name}: {bad.numel()} bad rows in {tiles.numel()} tiles; first tiles {tiles[:12].tolist()} (tile % 148: {[int(x) % 148 for x in tiles[:12]]}, tile // 148: {[int(x) // 148 for x in tiles[:12]]})", flush=True)
