import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "vaesne-dev_b200"), ROOT]
import torch
import model_cases as MC
from helpers import load_golden, golden_grads
orig = MC._check_grads
def verbose(model, g, tol=MC.GRAD_TOL, strip="", scale_floor=0.0):
    want = golden_grads(g)
    gmax = max(float(v.abs().max()) for v in want.values())
    rows = []
    for n, p in model.named_parameters():
        if not p.requires_grad: continue
        w = want[n]
        err = float((p.grad.cpu() - w).abs().max())
        rows.append((err / max(float(w.abs().max()), 1e-30), err / gmax, float(w.abs().max()) / gmax, n))
    rows.sort(reverse=True)
    print(f"--- tol {tol} floor {scale_floor}; gmax {gmax:.3e}")
    for r in sorted(rows, key=lambda r: -r[1])[:10]:
        print("  rel-own %.2e  rel-gmax %.2e  own/gmax %.2e  %s" % r)
    return (0.0, None)
MC._check_grads = verbose
MC.run_bright_case("bright_spec_elbo", "cuda")
