import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "vaesne-dev_b200"), ROOT]
import numpy as np, torch
import ops_cases as OC, attn_tc_ref as R
from VAESNe import _ops as P
dev = "cuda"
for case in [c for c in OC.ATTN_CASES_FULL if c["id"] in ("self_982_mask_rowmod", "self_983_masklen")]:
    (q, k, v, mask, mask_full, dO), (qd, kd, vd) = OC.make_attn_inputs(case, dev)
    seed_val, sid, p = 0x1234ABCD5678EF01, 77, 0.1
    seed = torch.tensor([seed_val], dtype=torch.int64, device=dev)
    N, Lq, Lk = case["N"], case["Lq"], case["Lk"]
    keep = R.keep_mask(seed_val, sid, p, N, 4, Lq, mask_full.numpy(), Lk)
    _, dscale = R.drop_threshold(p)
    o_ref, *_ = R.attn_reference_drop(q, k, v, mask_full, dO, torch.from_numpy(keep), dscale)
    O, LSE = P.attn_fwd(qd, kd, vd, mask.to(dev), P.Drop(p, seed, sid))
    err = (O.cpu().double() - o_ref).abs().view(N, Lq, 4, 8).amax(-1)    # [N, Lq, H]
    print(case["id"], "max", err.max().item(), "ref max", o_ref.abs().max().item())
    bad = (err > 1e-4).nonzero()
    print(" bad rows:", len(bad), "of", N * Lq * 4)
    print(bad[:40].tolist())
    kept = [int((~mask_full[n]).sum()) for n in range(N)]
    print(" kept keys per row", kept)
    # determinism + outliers without dropout
    O2, _ = P.attn_fwd(qd, kd, vd, mask.to(dev), P.Drop(p, seed, sid))
    print(" dropout run-to-run identical:", torch.equal(O, O2), (O - O2).abs().max().item())
    o0, *_ = OC.attn_reference(q, k, v, mask_full, dO)
    On, _ = P.attn_fwd(qd, kd, vd, mask.to(dev))
    On2, _ = P.attn_fwd(qd, kd, vd, mask.to(dev))
    e0 = (On.cpu().double() - o0).abs().view(N, Lq, 4, 8).amax(-1).flatten()
    print(" no-drop identical:", torch.equal(On, On2), "row err median", e0.median().item(), "top5", e0.topk(5).values.tolist())
    e1 = err.flatten()
    print(" drop row err median", e1.median().item(), "top5", e1.topk(5).values.tolist())
