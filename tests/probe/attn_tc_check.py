"""Dev probe (test infrastructure): accuracy + timing of the attention paths on the B200.
   VAESNE_NO_TC=1 python tests/probe/attn_tc_check.py   -> general kernels only"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "vaesne-dev_b200"), ROOT]
import torch
import ops_cases as OC
from helpers import rel_err
from VAESNe import _ops as P

dev = "cuda"
cases = [c for c in OC.ATTN_CASES_FULL if c["Lq"] >= 256 and c["Lk"] >= 256] + [
    dict(id="self_982_nomask", N=2, Lq=982, Lk=982, mask=False, packed="qkv"),
    dict(id="self_300x260_mask", N=3, Lq=300, Lk=260, mask=True, packed="q+kv"),
    dict(id="self_1024_mask", N=2, Lq=1024, Lk=1024, mask=True, packed="qkv"),
]
for scale in (() if os.environ.get('ONLY_TIME') else (1.0, 3.0)):
    for c in cases:
        (q, k, v, mask, mask_full, dO), (qd, kd, vd) = OC.make_attn_inputs(c, dev, scale)
        o_ref, lse_ref, dq_ref, dk_ref, dv_ref = OC.attn_reference(q, k, v, mask_full, dO)
        md = mask.to(dev) if mask is not None else None
        O, LSE = P.attn_fwd(qd, kd, vd, md)
        torch.cuda.synchronize()
        msg = f"scale={scale} {c['id']:28s} O {rel_err(O.cpu(), o_ref):.2e} LSE {rel_err(LSE.cpu(), lse_ref):.2e}"
        if os.environ.get("CHECK_BWD", "1") == "1":
            if c["packed"] == "qkv":
                dqkv = torch.zeros(c["N"], c["Lq"], 96, device=dev); dq, dk, dv = dqkv[..., :32], dqkv[..., 32:64], dqkv[..., 64:]
            else:
                dq = torch.zeros(c["N"], c["Lq"], 32, device=dev); dkv = torch.zeros(c["N"], c["Lk"], 64, device=dev); dk, dv = dkv[..., :32], dkv[..., 32:]
            P.attn_bwd(qd, kd, vd, md, O, LSE, dO.to(dev), dq, dk, dv)
            torch.cuda.synchronize()
            msg += f" dq {rel_err(dq.cpu(), dq_ref):.2e} dk {rel_err(dk.cpu(), dk_ref):.2e} dv {rel_err(dv.cpu(), dv_ref):.2e}"
        print(msg, flush=True)

import ctypes
if os.environ.get('TC_DBG'):
    ctypes.CDLL(os.path.join(ROOT, 'vaesne-dev_b200', 'lib', 'libvaesne_b200.so')).vaesne_debug_tc(int(os.environ['TC_DBG']))
# timing at the bench shape: N = 2*K*B = 1024 rows (B=64)
g = torch.Generator().manual_seed(0)
N = int(os.environ.get("TIME_N", "1024"))
qkv = torch.randn(N, 982, 96, generator=g).to(dev)
mask = (torch.rand(64, 982, generator=g) < 0.15).to(dev)
seed = torch.tensor([12345], dtype=torch.int64, device=dev)
for p in [float(x) for x in os.environ.get('PS', '0.0,0.1').split(',')]:
    drop = P.Drop(p, seed, 7) if p > 0 else P.NO_DROP
    for _ in range(2):
        O, LSE = P.attn_fwd(qkv[..., :32], qkv[..., 32:64], qkv[..., 64:], mask, drop)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        O, LSE = P.attn_fwd(qkv[..., :32], qkv[..., 32:64], qkv[..., 64:], mask, drop)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    el = N * 4 * 982 * 982
    print(f"fwd p={p}: {ms:.3f} ms  ({el / ms / 1e9:.2f} G score-elements/s nominal, {4 * 8 * el / ms / 1e9:.1f} TFLOP/s algorithmic)", flush=True)
    if os.environ.get("CHECK_BWD", "1") == "1":
        dO = torch.randn(N, 982, 32, generator=g).to(dev)
        dqkv = torch.empty(N, 982, 96, device=dev)
        for _ in range(2):
            P.attn_bwd(qkv[..., :32], qkv[..., 32:64], qkv[..., 64:], mask, O, LSE, dO, dqkv[..., :32], dqkv[..., 32:64], dqkv[..., 64:], drop)
        a.record()
        for _ in range(3):
            P.attn_bwd(qkv[..., :32], qkv[..., 32:64], qkv[..., 64:], mask, O, LSE, dO, dqkv[..., :32], dqkv[..., 32:64], dqkv[..., 64:], drop)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f"bwd p={p}: {ms:.3f} ms  ({10 * 8 * el / ms / 1e9:.1f} TFLOP/s algorithmic)", flush=True)

if os.environ.get('TC_PROF'):
    lib = ctypes.CDLL(os.path.join(ROOT, 'vaesne-dev_b200', 'lib', 'libvaesne_b200.so'))
    buf = (ctypes.c_longlong * 16)()
    torch.cuda.synchronize(); lib.vaesne_debug_tc_prof(buf)
    names = ["warp0 wait s_ready", "warp0 ld+compute (incl. in_free signal)", "warp0 wait::st", "warp0 wait o_ready", "warp0 total (after staging)", "warp0 wait out_free", "staging prologue", "warp0 wait dq product", "warp0 drain dq", "warp0 key-tile setup (to x_ready)", "warp0 key-tile tail (after o_ready)"]
    for n_, v in zip(names, buf):
        if n_: print(f"  bwd CTA(0,0) {n_:28s} {v:10d} clk")
